// Shared pieces of the second-generation tensor-core DyNCA kernels (dynca_tc2.cu forward, dynca_tc2_bwd.cu BPTT):
// tile / staging geometry, TMA + mbarrier + tcgen05 helpers, register-blocked perception, border patches, tensor maps.
#pragma once
#include <cuda.h>
#include <string.h>
#include <stdlib.h>
#include "dynca_tc_common.cuh"

#define T2_TH 8
#define T2_TW 16

// TMA (fp32, no swizzle) needs the innermost box coordinate to be a multiple of 4 elements (16 bytes; measured: any
// other start raises an illegal-instruction fault), so the boxes start 4 columns left of the tile: fine box columns
// x0-4 .. x0+19 (ring column x0-1 at index T2_XO), coarse box columns x0/2-4 .. x0/2+11 (footprint column x0/2-2 at T2_CO)
#define T2_XR 10
#define T2_XS 24
#define T2_XO 3
#define T2_CR 8
#define T2_CS 16
#define T2_CO 2
#define T2_QH 6
#define T2_QW 10
#define T2_THREADS 256
#define T2_HDR 1536u

// tile index -> (sample, y0, x0) without runtime integer division: q = floor(n / d) = umul64hi(n, ceil(2^64 / d)).
// The column is rotated by the row index (skew): a persistent CTA visits tiles blockIdx.x + k * gridDim.x, and with 16 tiles
// per row and 148 / 296 CTAs that stride is 4 / 8 columns, so without the rotation a quarter (an eighth) of the CTAs would
// process an image-border column every 4th (2nd) tile and the others never - border tiles cost ~40 % more, and the launch ends
// with the slowest CTA.  With the rotation every CTA cycles through all columns.
struct T2Tiles {
    int tiles_x, tiles_y, n_tiles, per_b, skew_mask;
    unsigned long long mx, mb;      // ceil(2^64 / tiles_x), ceil(2^64 / per_b); 0 when the divisor is 1
};
static inline T2Tiles t2_make_tiles(int B, int H, int W) {
    T2Tiles t;
    t.tiles_x = (W + T2_TW - 1) / T2_TW; t.tiles_y = (H + T2_TH - 1) / T2_TH;
    t.per_b = t.tiles_x * t.tiles_y; t.n_tiles = B * t.per_b;
    t.mx = t.tiles_x == 1 ? 0ull : ~0ull / (unsigned long long)t.tiles_x + 1ull;
    t.mb = t.per_b == 1 ? 0ull : ~0ull / (unsigned long long)t.per_b + 1ull;
    int p = 1;
    while (2 * p <= t.tiles_x) p *= 2;      // largest power of two <= tiles_x: the skew stays below tiles_x
    t.skew_mask = p - 1;
    return t;
}
__device__ __forceinline__ void t2_tile_decode(const T2Tiles& t, int tile, int& b, int& y0, int& x0) {
    const unsigned q1 = t.mx ? (unsigned)__umul64hi((unsigned long long)tile, t.mx) : (unsigned)tile;      // tile / tiles_x
    const unsigned bb = t.mb ? (unsigned)__umul64hi((unsigned long long)tile, t.mb) : (unsigned)tile;      // tile / per_b
    b = (int)bb;
    int tx = tile - (int)q1 * t.tiles_x + ((int)q1 & t.skew_mask);
    if (tx >= t.tiles_x) tx -= t.tiles_x;
    x0 = tx * T2_TW;
    y0 = ((int)q1 - (int)bb * t.tiles_y) * T2_TH;
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
        : "memory");
}
// 1-D bulk copies (TMA without a tensor map): shared -> global as a bulk group, global -> shared completing on an mbarrier
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t v[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t v[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t v[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// two floats -> bf16x2 with relu fused (lo in the low half)
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// packed fp32 pairs (FADD2 / FFMA2 on sm_100: one issue slot for two lanes of IEEE rn arithmetic - bit-identical to the scalar forms)
__device__ __forceinline__ float2 f2add(float2 a, float2 b) {
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) {       // a * b + c
    float2 r;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return f2fma(make_float2(-1.0f, -1.0f), b, a); }      // a - b (exact product)
__device__ __forceinline__ float2 f2fma2(float2 b, float2 c) { return f2fma(make_float2(2.0f, 2.0f), b, c); }      // 2 b + c

// 0xffff in each half whose bf16 value is non-zero (relu mask of a packed hidden pair: h != 0 <=> pre-activation > 0)
__device__ __forceinline__ uint32_t bf16x2_nz_mask(uint32_t h) {
    uint32_t m;
    asm("set.ne.u32.bf16x2 %0, %1, %2;" : "=r"(m) : "r"(h), "r"(0u));
    return m;
}

// perception of 4 vertically adjacent cells (rows r0 .. r0+3 of the tile) of one channel; col = stage column of the
// left neighbour.  Separable: s = [1 2 1]^T, d = [-1 0 1]^T over rows.
__device__ __forceinline__ void t2_percept4(const float* __restrict__ ch, int r0, int col, float id[4], float sx[4], float sy[4], float lp[4]) {
    // rows k = 0..3 as the pairs (0,1), (2,3) on packed fp32 (same operations and order as the scalar form, so bit-identical)
    float2 s[3][2], d[3][2], idp[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float v[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) v[k] = ch[(r0 + k) * T2_XS + T2_XO + col + j];
        const float2 v01 = make_float2(v[0], v[1]), v23 = make_float2(v[2], v[3]), v45 = make_float2(v[4], v[5]);
        const float2 v12 = make_float2(v[1], v[2]), v34 = make_float2(v[3], v[4]);
        s[j][0] = f2fma2(v12, f2add(v01, v23));
        s[j][1] = f2fma2(v34, f2add(v23, v45));
        d[j][0] = f2sub(v23, v01);
        d[j][1] = f2sub(v45, v23);
        if (j == 1) { idp[0] = v12; idp[1] = v34; }
    }
    const float2 m16 = make_float2(-16.0f, -16.0f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float2 sxp = f2sub(s[2][h], s[0][h]);
        const float2 syp = f2fma2(d[1][h], f2add(d[0][h], d[2][h]));
        const float2 lpp = f2fma(m16, idp[h], f2fma2(s[1][h], f2add(s[0][h], s[2][h])));
        id[2 * h] = idp[h].x; id[2 * h + 1] = idp[h].y;
        sx[2 * h] = sxp.x; sx[2 * h + 1] = sxp.y;
        sy[2 * h] = syp.x; sy[2 * h + 1] = syp.y;
        lp[2 * h] = lpp.x; lp[2 * h + 1] = lpp.y;
    }
}
// same for 3 vertically adjacent coarse cells (coarse stage stride T2_CS)
__device__ __forceinline__ void t2_percept3c(const float* __restrict__ ch, int r0, int col, float id[3], float sx[3], float sy[3], float lp[3]) {
    float s[3][3], d[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float v[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] = ch[(r0 + k) * T2_CS + T2_CO + col + j];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            s[j][k] = fmaf(2.0f, v[k + 1], v[k] + v[k + 2]);
            d[j][k] = v[k + 2] - v[k];
            if (j == 1) id[k] = v[k + 1];
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        sx[k] = s[2][k] - s[0][k];
        sy[k] = fmaf(2.0f, d[1][k], d[0][k] + d[2][k]);
        lp[k] = fmaf(-16.0f, id[k], fmaf(2.0f, s[1][k], s[0][k] + s[2][k]));
    }
}

// padding index map for coordinates at most one period outside the image (no generic modulo); ragged tiles reach
// further out only for cells nobody reads, so clamp
__device__ __forceinline__ int t2_padmap(int r, int n, int mode) {
    if (r >= 0 && r < n) return r;
    int v;
    if (mode == NCA_PAD_CIRCULAR) v = r < 0 ? r + n : r - n;
    else if (mode == NCA_PAD_REPLICATE) v = r;
    else v = r < 0 ? -r : 2 * (n - 1) - r;           // reflect
    return min(max(v, 0), n - 1);
}

// patch the staged tiles of a border tile: every staged position outside the image takes the value the padding mode
// prescribes (TMA filled it with zero, which is already right for constant padding).  One staged position per thread
// (fine: threads 0..179, coarse: threads 160..255), channel loop inside, so the index math runs once per tile.
template <int NS, int NT>
__device__ __forceinline__ void t2_patch_border(const DyncaGeom& g, const float* __restrict__ x, const float* __restrict__ xc, int b,
                                                int y0, int x0, float* __restrict__ sX, float* __restrict__ sXc) {
    const int C = g.C, H = g.H, W = g.W, tid = threadIdx.x;
    if (g.pad == NCA_PAD_CONSTANT) return;
    const bool wrap = g.pad == NCA_PAD_CIRCULAR;   // replicate / reflect sources lie inside the staged tile: smem -> smem
    if (tid < T2_XR * 18) {
        const int q = tid % 18, r = tid / 18;
        const int yy = y0 - 1 + r, xx = x0 - 1 + q;
        if (yy < 0 || yy >= H || xx < 0 || xx >= W) {
            const int iy = t2_padmap(yy, H, g.pad), ix = t2_padmap(xx, W, g.pad);
            float* dst = sX + r * T2_XS + T2_XO + q;
            if (wrap) {
                const size_t plane = (size_t)H * W;
                const float* src = x + (size_t)b * C * plane + (size_t)iy * W + ix;
                float v[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = c < C ? __ldg(src + c * plane) : 0.0f;
#pragma unroll
                for (int c = 0; c < 16; ++c) if (c < C) dst[c * T2_XR * T2_XS] = v[c];
            } else {
                const int sr = min(max(iy - (y0 - 1), 0), T2_XR - 1), sq = min(max(ix - (x0 - 1), 0), 17);
                const float* src = sX + sr * T2_XS + T2_XO + sq;
#pragma unroll 8
                for (int c = 0; c < C; ++c) dst[c * T2_XR * T2_XS] = src[c * T2_XR * T2_XS];
            }
        }
    }
    if (NS == 2 && tid >= NT - T2_CR * 12) {
        const int i = tid - (NT - T2_CR * 12);
        const int q = i % 12, r = i / 12;
        const int Hc = H >> 1, Wc = W >> 1;
        const int yy = (y0 >> 1) - 2 + r, xx = (x0 >> 1) - 2 + q;
        if (yy < 0 || yy >= Hc || xx < 0 || xx >= Wc) {
            const int iy = t2_padmap(yy, Hc, g.pad), ix = t2_padmap(xx, Wc, g.pad);
            float* dst = sXc + r * T2_CS + T2_CO + q;
            if (wrap) {
                const size_t cplane = (size_t)Hc * Wc;
                const float* src = xc + (size_t)b * C * cplane + (size_t)iy * Wc + ix;
                float v[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = c < C ? __ldg(src + c * cplane) : 0.0f;
#pragma unroll
                for (int c = 0; c < 16; ++c) if (c < C) dst[c * T2_CR * T2_CS] = v[c];
            } else {
                const int sr = min(max(iy - ((y0 >> 1) - 2), 0), T2_CR - 1), sq = min(max(ix - ((x0 >> 1) - 2), 0), 11);
                const float* src = sXc + sr * T2_CS + T2_CO + sq;
#pragma unroll 8
                for (int c = 0; c < C; ++c) dst[c * T2_CR * T2_CS] = src[c * T2_CR * T2_CS];
            }
        }
    }
}

// ---- programmatic dependent launch: the step kernels of a rollout are launched back to back on one stream; a launch with
// the programmatic-serialization attribute may become resident as soon as every CTA of its predecessor has executed
// griddep_launch (first instruction of the kernel) or exited, runs its prologue (operand images -> shared memory, barrier
// init, tensor-memory allocation) in the shadow of the predecessor's tail, and blocks in griddep_wait until the
// predecessor grid has completed and its memory is visible.  Without the attribute both are no-ops. ----
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t t2_launch(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    static int no_pdl = -1;                      // NCA_NO_PDL=1: plain stream order (debugging)
    if (no_pdl < 0) { const char* e = getenv("NCA_NO_PDL"); no_pdl = (e && e[0] == '1') ? 1 : 0; }
    cfg.attrs = at; cfg.numAttrs = (pdl && !no_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- synchronisation helpers: compute threads -> MMA warp hand-offs are mbarriers, not CTA barriers ----
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t p;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }   // named barrier over n threads
// D[tmem] (+)= A . B with precomputed 64-bit descriptors
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
}


// D[tmem] (+)= A[tmem] . B[smem]: the A operand (M = 128 rows = lanes, two bf16 K elements per 32-bit column) comes from
// tensor memory, so it costs no shared-memory bandwidth
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(db), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
}

// ---- compile-time specialisation: the step kernels are instantiated for the (C, fc) pairs of the reference's configurations
// (CT = FT = 0: any shape, geometry read from the arguments).  The kernel works on local copies of the geometry structs with the
// template values written in, so that every loop bound, shared-memory offset and descriptor that depends on C or fc folds to a
// constant (the generic instantiation spends ~15 % of its instructions on that arithmetic).
template <int CT, int FT>
__device__ __forceinline__ void t2_specialize(DyncaGeom& g, Bf16Geom& bg) {
    if (CT > 0) { g.C = CT; g.P = 4 * CT + g.cc; }
    if (FT > 0) g.fc = FT;
    if (CT > 0) {
        bg.npairs = (CT + 1) / 2;
        bg.K1 = ((bg.npairs + 1) * 8 + 15) / 16 * 16;
        bg.a1_bytes = (uint32_t)(bg.K1 / 8) * 2048u;
    }
    if (FT > 0) {
        bg.N1 = FT;
        bg.a2_bytes = (uint32_t)(FT / 8) * 2048u;
        bg.b2_bytes = (uint32_t)(FT / 8) * 256u;
    }
    if (CT > 0 && FT > 0) bg.b1_bytes = (uint32_t)(bg.K1 / 8) * (uint32_t)(FT / 8) * 128u;
}
static inline bool t2_nospec(const char* name) {      // NCA_T2_NOSPEC_FWD / NCA_T2_NOSPEC_BWD = 1: generic instantiation (debugging)
    const char* e = getenv(name);
    return e && e[0] == '1';
}
#define T2_DISPATCH_CF(C_, FC_, DO)                                  \
    do {                                                              \
        if (T2_NOSPEC) { DO(0, 0); }                                  \
        else if ((C_) == 16 && (FC_) == 128) { DO(16, 128); }         \
        else if ((C_) == 12 && (FC_) == 96) { DO(12, 96); }           \
        else if ((C_) == 13 && (FC_) == 96) { DO(13, 96); }           \
        else { DO(0, 0); }                                            \
    } while (0)

// ---- perception phases shared by the forward and the BPTT kernel (NW = number of compute warps, 8 or 16) ----
// fine perception -> A1: item = (channel pair, vertical block of 4 rows); lane = (column px, channel of the pair), channel
// in the low bit: channel planes are T2_XR * T2_XS = 240 floats apart = 16 banks, so the loads of a warp cover all 32 banks,
// and the 8-byte stores of a half warp (8 columns x 2 channels) are 128 contiguous bytes.
template <int NW>
__device__ __forceinline__ void t2_fine_to_a1(const float* __restrict__ sX, uint8_t* __restrict__ sA1, int C, int npairs, int warp, int lane) {
    const int hc = lane & 1, pxx = lane >> 1;
    for (int item = warp; item < 2 * npairs; item += NW) {
        const int cp = item >> 1, vb = item & 1, c = 2 * cp + hc;
        float id[4] = {0.f, 0.f, 0.f, 0.f}, sx[4] = {0.f, 0.f, 0.f, 0.f}, sy[4] = {0.f, 0.f, 0.f, 0.f}, lp[4] = {0.f, 0.f, 0.f, 0.f};
        if (c < C) t2_percept4(sX + c * T2_XR * T2_XS, 4 * vb, pxx, id, sx, sy, lp);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int rr = (4 * vb + k) * 16 + pxx;
            uint2 v;
            v.x = pack_bf16(id[k], sx[k]); v.y = pack_bf16(sy[k], lp[k]);
            *reinterpret_cast<uint2*>(sA1 + (uint32_t)cp * 2048u + (uint32_t)rr * 16u + (uint32_t)hc * 8u) = v;
        }
    }
}
// coarse perception -> Zc rows q = qy*10 + qx (64 rows per K chunk, chunk stride 1024): item = channel pair;
// block); lane = (qx, 3-row block).  Then the rows / chunks nobody writes are zeroed (they meet zero columns of U / zero rows of W1h
// but must be finite), and on border tiles the footprint is replicate-extended over the image border (edge clamp of
// the x2 bilinear upsample, dynca.py:93-94).  NT compute threads call this; bar_id = their named barrier.
template <int NW>
__device__ __forceinline__ void t2_coarse_to_zc(const DyncaGeom& g, const float* __restrict__ sXc, uint8_t* __restrict__ sZc, int npairs,
                                                int y0, int x0, bool border, int tid, int warp, int lane) {
    const int C = g.C;
    constexpr int NT = NW * 32;
    if (NW >= 16) {
        // one channel per warp (16 warps, at most 16 channels): half the dependent chain of the pair version below
        const int c = warp;
        if (lane < 2 * T2_QW && c < C) {
            const int hb = lane / T2_QW, qx = lane % T2_QW;
            float id0[3], sx0[3], sy0[3], lp0[3];
            t2_percept3c(sXc + c * T2_CR * T2_CS, 3 * hb, qx, id0, sx0, sy0, lp0);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int q = (3 * hb + k) * T2_QW + qx;
                uint2 v;
                v.x = pack_bf16(id0[k], sx0[k]); v.y = pack_bf16(sy0[k], lp0[k]);
                *reinterpret_cast<uint2*>(sZc + (uint32_t)(c >> 1) * 1024u + (uint32_t)q * 16u + (uint32_t)(c & 1) * 8u) = v;
            }
        }
        if ((C & 1) && warp == C && lane < 2 * T2_QW) {       // odd channel count: the upper half of the last pair is zero
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int q = (3 * (lane / T2_QW) + k) * T2_QW + lane % T2_QW;
                *reinterpret_cast<uint2*>(sZc + (uint32_t)(C >> 1) * 1024u + (uint32_t)q * 16u + 8u) = make_uint2(0u, 0u);
            }
        }
    } else if (lane < 2 * T2_QW) {
        for (int cp = warp; cp < npairs; cp += NW) {
            const int hb = lane / T2_QW, qx = lane % T2_QW;
            float id0[3], sx0[3], sy0[3], lp0[3], id1[3] = {0.f, 0.f, 0.f}, sx1[3] = {0.f, 0.f, 0.f}, sy1[3] = {0.f, 0.f, 0.f}, lp1[3] = {0.f, 0.f, 0.f};
            t2_percept3c(sXc + (2 * cp) * T2_CR * T2_CS, 3 * hb, qx, id0, sx0, sy0, lp0);
            if (2 * cp + 1 < C) t2_percept3c(sXc + (2 * cp + 1) * T2_CR * T2_CS, 3 * hb, qx, id1, sx1, sy1, lp1);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int q = (3 * hb + k) * T2_QW + qx;
                uint4 v;
                v.x = pack_bf16(id0[k], sx0[k]); v.y = pack_bf16(sy0[k], lp0[k]);
                v.z = pack_bf16(id1[k], sx1[k]); v.w = pack_bf16(sy1[k], lp1[k]);
                *reinterpret_cast<uint4*>(sZc + (uint32_t)cp * 1024u + (uint32_t)q * 16u) = v;
            }
        }
    }
    if (tid < 32) *reinterpret_cast<uint4*>(sZc + (uint32_t)(tid >> 2) * 1024u + (uint32_t)(60 + (tid & 3)) * 16u) = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < (8 - npairs) * 64; i += NT)
        *reinterpret_cast<uint4*>(sZc + (uint32_t)(npairs + (i >> 6)) * 1024u + (uint32_t)(i & 63) * 16u) = make_uint4(0, 0, 0, 0);
    if (border) {
        bar_sync_n(1, NT);
        const int Hc = g.H >> 1, Wc = g.W >> 1;
        for (int i = tid; i < T2_QH * T2_QW * 8; i += NT) {
            const int ch = i & 7, q = i >> 3;
            const int qy = q / T2_QW, qxx = q % T2_QW;
            const int Qy = (y0 >> 1) - 1 + qy, Qx = (x0 >> 1) - 1 + qxx;
            const int Cy = min(max(Qy, 0), Hc - 1), Cx = min(max(Qx, 0), Wc - 1);
            if (Cy == Qy && Cx == Qx) continue;
            // a ragged last tile can clamp to a cell outside the 6x10 footprint only if that cell is outside every
            // in-image fine cell's support; keep the index in range
            const int sy_ = min(max(Cy - ((y0 >> 1) - 1), 0), T2_QH - 1), sx_ = min(max(Cx - ((x0 >> 1) - 1), 0), T2_QW - 1);
            const int qs = sy_ * T2_QW + sx_;
            *reinterpret_cast<uint4*>(sZc + (uint32_t)ch * 1024u + (uint32_t)q * 16u) =
                *reinterpret_cast<const uint4*>(sZc + (uint32_t)ch * 1024u + (uint32_t)qs * 16u);
        }
    }
}
// cond chunk of A1 (see dynca_cond_chunk) from the TMA-staged cond tile sCond [cc][8][16] (NCA_COND_TENSOR)
__device__ __forceinline__ uint4 t2_cond_chunk_smem(const DyncaGeom& g, const float* __restrict__ sCond, int r, bool inimg) {
    float cv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (inimg) {
        const bool split = dynca_cond_split(g.cc);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int src = i < g.cc ? i : ((split && i >= g.cc + 2 && i < 2 * g.cc + 2) ? i - g.cc - 2 : -1);
            if (src >= 0) {
                const float raw = sCond[src * (T2_TH * T2_TW) + r];
                const float hi = __bfloat162float(__float2bfloat16_rn(raw));
                cv[i] = i < g.cc ? hi : raw - hi;
            } else if (i == g.cc || i == g.cc + 1) cv[i] = 1.0f;
        }
    }
    uint4 v;
    v.x = pack_bf16(cv[0], cv[1]); v.y = pack_bf16(cv[2], cv[3]); v.z = pack_bf16(cv[4], cv[5]); v.w = pack_bf16(cv[6], cv[7]);
    return v;
}
// fire decisions of one 8x16 tile by ONE warp: 32 lanes = 32 quads of 4 consecutive pixels -> sFire[128]
__device__ __forceinline__ void t2_fire_tile(const FireMask& fm, int b, int y0, int x0, int H, int W, int lane, float* __restrict__ sFire, int enc = 0) {
    const int fy = y0 + (lane >> 2), fx = x0 + 4 * (lane & 3);
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    if (fy < H && fx < W) {
        const uint32_t p = (uint32_t)(fy * W + fx);
        const uint4 rr = nca_philox4x32_10(p >> 2, (uint32_t)b, fm.t, NCA_PHILOX_STREAM, fm.k0, fm.k1);
        f.x = nca_fire(rr.x, fm.thr, enc); f.y = nca_fire(rr.y, fm.thr, enc);
        f.z = nca_fire(rr.z, fm.thr, enc); f.w = nca_fire(rr.w, fm.thr, enc);
    }
    *reinterpret_cast<float4*>(sFire + (lane >> 2) * 16 + 4 * (lane & 3)) = f;
}

// ---- host side: tensor maps ---------------------------------------------------------------------------------
typedef CUresult (*T2EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline T2EncodeFn t2_encode_fn() {
    static T2EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (T2EncodeFn)p;
    }
    return fn;
}
// 5-D map over [slots][B][C][Hh][Ww] fp32 with box [1][1][C][bh][bw]
static inline int t2_make_map(CUtensorMap* tm, const float* base, int slots, size_t slot_floats, int B, int C, int Hh, int Ww, int bh, int bw, int bc = 0) {
    T2EncodeFn fn = t2_encode_fn();
    if (!fn) { nca_set_error("cuTensorMapEncodeTiled is not available from the driver"); return NCA_ERR_CUDA; }
    cuuint64_t dims[5] = {(cuuint64_t)Ww, (cuuint64_t)Hh, (cuuint64_t)C, (cuuint64_t)B, (cuuint64_t)slots};
    cuuint64_t strides[4] = {(cuuint64_t)Ww * 4, (cuuint64_t)Hh * Ww * 4, (cuuint64_t)C * Hh * Ww * 4, (cuuint64_t)slot_floats * 4};
    cuuint32_t box[5] = {(cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)(bc > 0 ? bc : C), 1, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult rc = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { nca_set_error("cuTensorMapEncodeTiled failed with %d", (int)rc); return NCA_ERR_CUDA; }
    return NCA_OK;
}

// CTAs of a kernel that are co-resident on an SM by registers (the occupancy API answers 1 for the kernels that allocate tensor
// memory, so count by hand: registers are allocated per warp in units of 256)
template <typename K>
static inline int t2_occupancy_by_regs(K kernel, int threads) {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess || fa.numRegs <= 0) return 1;
    const int per_warp = (fa.numRegs * 32 + 255) / 256 * 256, warps = (threads + 31) / 32;
    const int occ = 65536 / (per_warp * warps);
    return occ < 1 ? 1 : occ;
}

static inline int t2_num_sms() { return nca_sm_count(); }

