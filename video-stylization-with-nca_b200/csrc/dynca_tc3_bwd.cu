// DyNCA BPTT step on the 5th-gen tensor cores, third generation (companion of dynca_tc2.cu; replaces dynca_tc2_bwd.cu).
// Replaces autograd's replay of ExtraChannels/models/dynca.py:113-123 for one step:
//   given the perception operand Z_t the forward recorded (operand history) and g = dL/dx_{t+1}  ->  dL/dx_t and the weight gradients.
//
// What bounded the second generation was the shared-memory pipe (operand fetch of the MMAs + the fp32 planes of the transposed
// stencils: ~7.3 k of the 8.3 k cycles of a tile).  Here the gradient of the perception vector is produced TRANSPOSED,
//   D6^T [k' = 4c + filter][cell] = W1h^T . Ga^T
// so that a TMEM lane holds one (channel, filter) plane of the 8x16 tile, a row of 16 cells per 16 columns: the transposed 3x3
// stencils run in REGISTERS straight out of tensor memory (separable, per-lane coefficient vectors), the four filter lanes of a
// channel are summed by a transposing shuffle reduction (14 shuffles per row of 18 outputs) and go out as red.global.add.v4.
// No fp32 planes, no shared-memory traffic and no CTA barrier in the stencil phases.  Rows 64..127 of the A operand repeat rows
// 0..63, so all four lane quarters hold the planes: quarters 0 / 1 process tile rows 0..3, quarters 2 / 3 rows 4..7.
// Two perception scales: the forward folds the coarse scale into the operand (Z = fine + U . Zc, dynca_tc2.cu), so the recompute is
// ONE GEMM and the weight gradient pairs Ga with Z; the coarse part of the state gradient is D7^T = bf16(D6^T) . U (A operand from
// tensor memory, written in place by the stencil warps), then the same register stencil on the 6x10 coarse footprint -> red.add
// into the coarse gradient buffer (its 2x2-mean transpose is applied when the next step reads g, as before).
//
// Warp roles (17 warps, 1 CTA / SM, persistent over tiles):
//   E warps 0..7  : g tile -> gn = g + 0.25 gc (+ tap), Gy = fire * gn (bf16); the residual path dL/dx_t += gn as one red.add.v4 per
//                   four cells; E1: H = relu(D1), Ga = D3 * [D1 > 0] -> bf16 operands in shared memory; they also issue the TMA /
//                   bulk loads of the next tiles
//   S warps 8..15 : two groups of four (one warp per lane quarter) that alternate tiles: fine stencil, bf16(D6^T) in place, coarse stencil
//   warp 16       : MMA issue.  D1 = Z.W1h^T, D3 = Gy.W2 | D6^T, D4 += H^T.Gy, D5 += Ga^T.Z | D7^T
// TMEM (480 of 512 columns): D1 128 | region 0 128 | region 1 128 | D4 16 | D5 80; a tile uses region (tile & 1) for D3, then for
// D6^T, whose first 64 columns are overwritten by the packed bf16 copy and whose last 64 take D7^T - so E1 of tile i+1 (D3 in the
// other region) overlaps the stencils of tile i.
#include "dynca_tc2.cuh"

#define T3_NTHREADS 544
#define T3_NE 256          // threads of the E warps
#define T3_HDR 4096u
#define T3_D1 0u
#define T3_R0 128u
#define T3_D4 384u
#define T3_D5 400u

struct T3BwdArgs {
    DyncaGeom g;
    Bf16Geom bg;
    float* g_in; float* gc_in;                        // dL/dx_{t+1} and its coarse part (read through TMA; zeroed if zero_in)
    int zero_in, zero_cin;
    const float* g_tap; int tap_c; float tap_scale;   // optional rgb tap at states[t+1]
    float* g_out; float* gc_out;                      // dL/dx_t (red.add) and its coarse part (red.add)
    const __nv_bfloat16* B1; const __nv_bfloat16* B2d; const __nv_bfloat16* U; const __nv_bfloat16* A6;
    float* gW1p; float* gW2p; float* gb2p;            // fp32 accumulators, padded fp32-path layout (red.add)
    FireMask fm;
    T2Tiles tl;
    const uint8_t* op_in;                             // operand history of this step: Z per tile
    long long* tdbg;                                  // optional phase timestamps of CTA 0 (debug, -DNCA_T2_TIMING)
};
#ifdef NCA_T2_TIMING
#define T3_STAMP(it_, k_) do { if (a.tdbg && blockIdx.x == 0 && (it_) < 8) a.tdbg[(it_) * 32 + (k_)] = clock64(); } while (0)
#else
#define T3_STAMP(it_, k_) do { } while (0)
#endif

struct T3Smem { uint32_t b1, b2d, u, a6, z, gy, h, ga, gn, gcn, total; };
__host__ __device__ static inline T3Smem t3_smem(const DyncaGeom& g, const Bf16Geom& bg) {
    T3Smem s;
    const uint32_t C = (uint32_t)g.C;
    uint32_t o = T3_HDR;
    s.b1 = o; o += bg.b1_bytes;
    s.b2d = o; o += (uint32_t)(g.fc / 8) * 256u;
    o = (o + 127u) & ~127u;
    s.u = o; o += g.ns == 2 ? 16384u : 0u;
    s.a6 = o; o += (uint32_t)(g.fc / 8) * 2048u;
    s.z = o; o += 2u * bg.a1_bytes;
    s.gy = o; o += 2u * 4096u;
    s.h = o; o += 16u * 2048u;        // 16 chunks whatever fc is: the MN-major M = 128 views read all of them
    s.ga = o; o += 16u * 2048u;
    s.gn = o; o += C * T2_TH * T2_TW * 4u;
    o = (o + 127u) & ~127u;
    s.gcn = o; o += g.ns == 2 ? C * 32u * 4u : 0u;
    o = (o + 127u) & ~127u;
    s.total = o;
    return s;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {      // p 16-byte aligned
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {                        // p 8-byte aligned
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t v[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
// wait for the tensor-memory loads; the registers of the last load are operands, so that no use of them can be scheduled above
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}
// image index that padded coordinate r of an axis of n cells folds back to under the TRANSPOSED padding (-1 = dropped)
__device__ __forceinline__ int t3_fold(int r, int n, int mode) {
    if (r >= 0 && r < n) return r;
    if (mode == NCA_PAD_CONSTANT || r < -1 || r > n) return -1;      // only the 1-cell ring of the image carries gradient
    if (mode == NCA_PAD_CIRCULAR) return r < 0 ? n - 1 : 0;
    if (mode == NCA_PAD_REPLICATE) return r < 0 ? 0 : n - 1;
    return r < 0 ? 1 : n - 2;                                         // reflect
}

// per-lane description of the transposed stencil of filter f (k' = 4c + f: identity, Sobel-x, Sobel-y, Laplacian).  The forward
// filter is K[a][b] = v[a] h[b] (+ ctr at the centre), so the gradient at (y', x') is sum_a v[a] sum_b h[b] G(y' - a + 1, x' - b + 1):
// T(y, p) = h0 G[p] + h1 G[p-1] + h2 G[p-2] for the ring position p = x' + 1, and input row y adds v0 T to output row y - 1,
// v1 T (+ ctr G) to row y and v2 T to row y + 1.
struct T3Coef { float h0, h1, h2, v0, v1, v2, ctr; };
__device__ __forceinline__ T3Coef t3_coef(int f) {
    T3Coef k;
    k.h0 = f == 0 ? 0.f : (f == 1 ? -1.f : 1.f);
    k.h1 = f == 0 ? 1.f : (f == 1 ? 0.f : 2.f);
    k.h2 = f == 0 ? 0.f : 1.f;
    k.v0 = f == 0 ? 0.f : (f == 2 ? -1.f : 1.f);
    k.v1 = f == 0 ? 1.f : (f == 2 ? 0.f : 2.f);
    k.v2 = f == 0 ? 0.f : 1.f;
    k.ctr = f == 3 ? -16.f : 0.f;
    return k;
}

// transposing reduction of one output row of 18 ring positions over the four filter lanes of a channel (lane bits 0 / 1):
// afterwards lane (b1, b0) holds the sums of positions (0,0): 0..4   (1,0): 5..8   (0,1): 17..13   (1,1): 12..9   in u[0..]
__device__ __forceinline__ void t3_reduce18(const float (&o)[18], bool b0, bool b1, float (&u)[5]) {
    float w[10];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const float send = b0 ? o[i] : o[17 - i];
        const float keep = b0 ? o[17 - i] : o[i];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    w[9] = 0.0f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const float send = b1 ? w[i] : w[5 + i];
        const float keep = b1 ? w[5 + i] : w[i];
        u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
}
// same for 12 ring positions of the coarse footprint; lane (b1, b0) ends with (0,0): 0, 1   (1,0): 2..5   (0,1): 11, 10   (1,1): 9..6
__device__ __forceinline__ void t3_reduce12(const float (&o)[12], bool b0, bool b1, float (&u)[4]) {
    float w[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float send = b0 ? o[i] : o[11 - i];
        const float keep = b0 ? o[11 - i] : o[i];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    // b1 = 0 keeps w[0..1] and receives the partner's w[0..1]; b1 = 1 keeps w[2..5] and receives the partner's w[2..5]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b1 ? w[i < 2 ? i : 0] : w[2 + i];
        const float keep = b1 ? w[2 + i] : w[i < 2 ? i : 0];
        const float r = __shfl_xor_sync(0xffffffffu, send, 2);
        u[i] = keep + r;
    }
}

struct T3Emit {
    float* gob;          // g_out + (b * C + c) * plane
    int y0, x0, H, W, pad;
    bool fast;           // every ring position of the tile is inside the image
    bool b0, b1, act;
};
// one fine output row (tile-relative row oy = -1 .. 8) -> red.add into dL/dx_t
__device__ __forceinline__ void t3_emit_fine(const T3Emit& e, int oy, const float (&o)[18]) {
    float u[5];
    t3_reduce18(o, e.b0, e.b1, u);
    const int gy = e.y0 + oy;
    if (e.fast) {
        if (e.act) {
            float* rowp = e.gob + (size_t)gy * e.W + e.x0;
            // the four cells of this lane in image order, and their column offset
            const float c0 = e.b1 ? u[0] : u[1], c1 = e.b1 ? u[1] : u[2], c2 = e.b1 ? u[2] : u[3], c3 = e.b1 ? u[3] : u[4];
            const float a0 = e.b0 ? c3 : c0, a1 = e.b0 ? c2 : c1, a2 = e.b0 ? c1 : c2, a3 = e.b0 ? c0 : c3;
            const int xo = e.b0 ? (e.b1 ? 8 : 12) : (e.b1 ? 4 : 0);
            red_add_v4(rowp + xo, a0, a1, a2, a3);
            if (!e.b1) atomicAdd(rowp + (e.b0 ? T2_TW : -1), u[0]);
        }
    } else if (e.act) {
        const int ty = t3_fold(gy, e.H, e.pad);
        const int n = e.b1 ? 4 : 5, pb = e.b1 ? (e.b0 ? 12 : 5) : (e.b0 ? 17 : 0), d = e.b0 ? -1 : 1;
        if (ty >= 0) {
#pragma unroll 1
            for (int i = 0; i < n; ++i) {
                const int tx = t3_fold(e.x0 - 1 + pb + d * i, e.W, e.pad);
                const float v = i == 0 ? u[0] : (i == 1 ? u[1] : (i == 2 ? u[2] : (i == 3 ? u[3] : u[4])));
                if (tx >= 0) atomicAdd(e.gob + (size_t)ty * e.W + tx, v);
            }
        }
    }
}
// one coarse output row: coarse image row gy, ring position p <-> coarse column cxm1 + p
__device__ __forceinline__ void t3_emit_coarse(const T3Emit& e, float* gcb, int gy, int cxm1, int Hc, int Wc, const float (&o)[12]) {
    float u[4];
    t3_reduce12(o, e.b0, e.b1, u);
    if (e.fast) {
        if (e.act) {
            float* rowp = gcb + (size_t)gy * Wc + cxm1;
            const float a0 = e.b0 ? u[3] : u[0], a1 = e.b0 ? u[2] : u[1], a2 = e.b0 ? u[1] : u[2], a3 = e.b0 ? u[0] : u[3];
            if (e.b1) red_add_v4(rowp + (e.b0 ? 6 : 2), a0, a1, a2, a3);
            else red_add_v2(rowp + (e.b0 ? 10 : 0), e.b0 ? u[1] : u[0], e.b0 ? u[0] : u[1]);
        }
    } else if (e.act) {
        const int ty = t3_fold(gy, Hc, e.pad);
        const int n = e.b1 ? 4 : 2, pb = e.b1 ? (e.b0 ? 9 : 2) : (e.b0 ? 11 : 0), d = e.b0 ? -1 : 1;
        if (ty >= 0) {
#pragma unroll 1
            for (int i = 0; i < n; ++i) {
                const int tx = t3_fold(cxm1 + pb + d * i, Wc, e.pad);
                const float v = i == 0 ? u[0] : (i == 1 ? u[1] : (i == 2 ? u[2] : u[3]));
                if (tx >= 0) atomicAdd(gcb + (size_t)ty * Wc + tx, v);
            }
        }
    }
}

template <int NS, int CT, int FT>
__global__ void __launch_bounds__(T3_NTHREADS, 1) dynca_bwd_tc3_kernel(const __grid_constant__ CUtensorMap tm_g,
                                                                        const __grid_constant__ CUtensorMap tm_gc, const T3BwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    DyncaGeom g = a.g;
    Bf16Geom bg = a.bg;
    t2_specialize<CT, FT>(g, bg);
    const T3Smem L = t3_smem(g, bg);
    uint64_t* barT = reinterpret_cast<uint64_t*>(smem);            // TMA: g tile loaded
    uint64_t* barO = reinterpret_cast<uint64_t*>(smem + 8);        // [2] Z of buffer 0 / 1 loaded
    uint64_t* barG = reinterpret_cast<uint64_t*>(smem + 24);       // [2] E -> MMA: Gy / gn images written (256)
    uint64_t* barM2 = reinterpret_cast<uint64_t*>(smem + 40);      // MMA -> E: D1, D3
    uint64_t* barC = reinterpret_cast<uint64_t*>(smem + 48);       // E -> MMA: H, Ga written, D1 / D3 consumed (256)
    uint64_t* barM3a = reinterpret_cast<uint64_t*>(smem + 56);     // [2] MMA -> S group: D6^T
    uint64_t* barM3b = reinterpret_cast<uint64_t*>(smem + 72);     // MMA -> E: D4, D5 (operand buffers free)
    uint64_t* barD = reinterpret_cast<uint64_t*>(smem + 80);       // [2] S group -> MMA: D6^T consumed, packed copy written (128)
    uint64_t* barM4 = reinterpret_cast<uint64_t*>(smem + 96);      // [2] MMA -> S group: D7^T
    uint64_t* barE = reinterpret_cast<uint64_t*>(smem + 112);      // [2] S group -> MMA: D7^T consumed (128)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 128);
    uint64_t* barF = reinterpret_cast<uint64_t*>(smem + 192);      // [4] MMA warp -> E: fire table of a tile written
    float* sFire4 = reinterpret_cast<float*>(smem + 256);          // 4 x 128 floats (ring over tiles)
    uint8_t* sB1 = smem + L.b1;
    uint8_t* sB2d = smem + L.b2d;
    uint8_t* sU = smem + L.u;
    uint8_t* sA6 = smem + L.a6;
    uint8_t* sZ2 = smem + L.z;            // 2 x a1_bytes
    uint8_t* sGy2 = smem + L.gy;          // 2 x 4096
    uint8_t* sH = smem + L.h;
    uint8_t* sGa = smem + L.ga;
    float* sGn = reinterpret_cast<float*>(smem + L.gn);
    float* sGcn = reinterpret_cast<float*>(smem + L.gcn);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = g.C, H = g.H, W = g.W, fc = g.fc;
    const size_t plane = (size_t)H * W;
    const int n_tiles = a.tl.n_tiles;
    const int n_my = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // tiles of this CTA (>= 1)
    const uint32_t stage_bytes = (uint32_t)C * (T2_TH * T2_TW + (NS == 2 ? 32 : 0)) * 4u;
    const uint32_t op_bytes = bg.a1_bytes;

    griddep_launch();
    // ---- one-time setup (independent of the previous launch's output: may overlap its tail) ----
    for (uint32_t i = tid; i < bg.b1_bytes / 16; i += T3_NTHREADS)
        reinterpret_cast<uint4*>(sB1)[i] = __ldg(reinterpret_cast<const uint4*>(a.B1) + i);
    for (uint32_t i = tid; i < (uint32_t)(fc / 8) * 256u / 16; i += T3_NTHREADS)
        reinterpret_cast<uint4*>(sB2d)[i] = __ldg(reinterpret_cast<const uint4*>(a.B2d) + i);
    if (NS == 2)
        for (uint32_t i = tid; i < 16384u / 16; i += T3_NTHREADS)
            reinterpret_cast<uint4*>(sU)[i] = __ldg(reinterpret_cast<const uint4*>(a.U) + i);
    for (uint32_t i = tid; i < (uint32_t)(fc / 8) * 2048u / 16; i += T3_NTHREADS)
        reinterpret_cast<uint4*>(sA6)[i] = __ldg(reinterpret_cast<const uint4*>(a.A6) + i);
    // everything an MMA may read before the tile loop writes it must be finite: clear the operand buffers once
    for (uint32_t i = L.z / 16 + tid; i < L.gn / 16; i += T3_NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(barT, 1);
        mbar_init(barO, 1); mbar_init(barO + 1, 1);
        mbar_init(barG, T3_NE); mbar_init(barG + 1, T3_NE);
        mbar_init(barM2, 1); mbar_init(barC, T3_NE);
        mbar_init(barM3a, 1); mbar_init(barM3a + 1, 1); mbar_init(barM3b, 1);
        mbar_init(barD, 128); mbar_init(barD + 1, 128);
        mbar_init(barM4, 1); mbar_init(barM4 + 1, 1);
        mbar_init(barE, 128); mbar_init(barE + 1, 128);
        mbar_init(barF, 1); mbar_init(barF + 1, 1); mbar_init(barF + 2, 1); mbar_init(barF + 3, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) tmem_alloc(tmem_slot, 512u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();          // dL/dx_{t+1} (and the weight-gradient accumulators) of the previous launch are complete from here on

    if (warp == 16) {
        // =========================== MMA warp ===========================
        // geometry of this warp deliberately from the ARGUMENTS (see dynca_tc2_bwd.cu history: with every descriptor a compile-time
        // constant nvcc 12.9 mis-generates the two-scale (16, 128) instantiation)
        const uint32_t lbo_b1 = (uint32_t)(a.g.fc / 8) * 128u;
        const uint32_t id_fc = umma_idesc_bf16(128, fc);
        const uint32_t id_w2 = umma_idesc_bf16_mn(128, 16), id_w1 = umma_idesc_bf16_mn(128, bg.K1);
        const uint32_t id_t = umma_idesc_bf16(128, 128);
        const uint32_t id_c = umma_idesc_bf16(128, 64) | (1u << 16);                   // B operand (U) MN-major
        const uint64_t dB1 = umma_desc(smem_u32(sB1), lbo_b1, 128u);
        const uint64_t dB2d = umma_desc(smem_u32(sB2d), lbo_b1, 128u);
        const uint64_t dA6 = umma_desc(smem_u32(sA6), 2048u, 128u);
        const uint64_t dGa = umma_desc(smem_u32(sGa), 2048u, 128u);                    // Ga as [N = cell][K = hidden], K-major
        const uint64_t dHt = umma_desc(smem_u32(sH), 128u, 2048u), dGat = umma_desc(smem_u32(sGa), 128u, 2048u);      // MN-major views
        const uint64_t dUb = umma_desc(smem_u32(sU), 128u, 2048u);                     // U as [K = cell][N = coarse cell], MN-major B
        const uint64_t dZ_0 = umma_desc(smem_u32(sZ2), 2048u, 128u), dZt_0 = umma_desc(smem_u32(sZ2), 128u, 2048u);
        const uint64_t dGy_0 = umma_desc(smem_u32(sGy2), 2048u, 128u), dGyt_0 = umma_desc(smem_u32(sGy2), 128u, 2048u);
        const uint64_t oZ = (uint64_t)(bg.a1_bytes >> 4), oGy = (uint64_t)(4096u >> 4);
        const uint64_t sB1k = (uint64_t)((2u * lbo_b1) >> 4);
        const int k1steps = bg.K1 / 16, kfsteps = fc / 16;
        const bool leader = elect_one();
        auto recompute = [&](int it) {      // D1 = Z . W1h^T, D3 = Gy . W2 (into the region of the tile)
            const uint64_t par = (uint64_t)(it & 1);
            const uint32_t ph = (uint32_t)((it >> 1) & 1);
            mbar_wait(barO + (it & 1), ph);
            mbar_wait(barG + (it & 1), ph);
            tc_fence_after();
            if (leader) {
#pragma unroll 5
                for (int ks = 0; ks < k1steps; ++ks)
                    umma_ss(tmem_base + T3_D1, dZ_0 + par * oZ + (uint64_t)(ks * (4096 >> 4)), dB1 + (uint64_t)ks * sB1k, id_fc, ks > 0);
                umma_ss(tmem_base + T3_R0 + 128u * (uint32_t)par, dGy_0 + par * oGy, dB2d, id_fc, false);
                umma_commit(barM2);
            }
        };
        auto coarse = [&](int it) {         // D7^T = bf16(D6^T) . U, A operand from tensor memory
            const uint32_t R = tmem_base + T3_R0 + 128u * (uint32_t)(it & 1);
            mbar_wait(barD + (it & 1), (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_ts(R + 64u, R + 8u * (uint32_t)ks, dUb + (uint64_t)(ks * (256 >> 4)), id_c, ks > 0);
                umma_commit(barM4 + (it & 1));
            }
        };
        // fire decisions of a tile (Philox, one quad per lane) in the idle time of this warp, three tiles ahead of their use
        auto fire_table = [&](int it) {
            if (a.fm.supplied || it >= n_my) return;
            int tb_, ty_, tx_;
            t2_tile_decode(a.tl, (int)blockIdx.x + it * (int)gridDim.x, tb_, ty_, tx_);
            t2_fire_tile(a.fm, tb_, ty_, tx_, H, W, lane, sFire4 + (it & 3) * 128);
            __syncwarp();
            if (lane == 0) mbar_arrive(barF + (it & 3));
        };
        fire_table(0);
        fire_table(1);
        fire_table(2);
        recompute(0);
        for (int it = 0; it < n_my; ++it) {
            const uint64_t par = (uint64_t)(it & 1);
            const uint32_t R = tmem_base + T3_R0 + 128u * (uint32_t)par;
            mbar_wait(barC, (uint32_t)(it & 1));               // H, Ga written; D1 / D3 consumed
            tc_fence_after();
            if (leader) T3_STAMP(it, 16);
            if (leader) {
                // g_z^T first (the stencil warps wait for it), weight gradients behind it
#pragma unroll 8
                for (int ks = 0; ks < kfsteps; ++ks)           // D6^T = W1h^T . Ga^T
                    umma_ss(R, dA6 + (uint64_t)(ks * (4096 >> 4)), dGa + (uint64_t)(ks * (4096 >> 4)), id_t, ks > 0);
                umma_commit(barM3a + (it & 1));
                const uint64_t dGyt = dGyt_0 + par * oGy, dZt = dZt_0 + par * oZ;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {               // 16 cells per instruction
                    const uint64_t o = (uint64_t)(ks * (256 >> 4));
                    umma_ss(tmem_base + T3_D4, dHt + o, dGyt + o, id_w2, !(it == 0 && ks == 0));
                    umma_ss(tmem_base + T3_D5, dGat + o, dZt + o, id_w1, !(it == 0 && ks == 0));
                }
                umma_commit(barM3b);
            }
            if (leader) T3_STAMP(it, 17);
            if (NS == 2 && it >= 1) coarse(it - 1);
            if (leader) T3_STAMP(it, 18);
            if (it + 1 < n_my) {
                if (it >= 1) {                                 // the region of tile it + 1 was the region of tile it - 1
                    mbar_wait((NS == 2 ? barE : barD) + ((it - 1) & 1), (uint32_t)(((it - 1) >> 1) & 1));
                    tc_fence_after();
                }
                recompute(it + 1);
            }
            fire_table(it + 3);
            if (leader) T3_STAMP(it, 19);
        }
        if (NS == 2) coarse(n_my - 1);
    } else if (warp < 8) {
        // =========================== E warps ===========================
        const int r = tid & 127, eh = tid >> 7;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t row_off = (uint32_t)r * 16u;
        const int py = r >> 4, px = r & 15;
        const CUtensorMap* const ptm_g = &tm_g;
        const CUtensorMap* const ptm_gc = &tm_gc;
        float b2acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        auto issue_g = [&](int it) {        // one thread: TMA of the g tile (and its coarse part) of local tile it
            int tb_, ty_, tx_;
            t2_tile_decode(a.tl, (int)blockIdx.x + it * (int)gridDim.x, tb_, ty_, tx_);
            mbar_expect_tx(barT, stage_bytes);
            tma_load_5d(sGn, ptm_g, barT, tx_, ty_, 0, tb_, 0);
            if (NS == 2) tma_load_5d(sGcn, ptm_gc, barT, tx_ >> 1, ty_ >> 1, 0, tb_, 0);
        };
        auto issue_z = [&](int it) {        // one thread: bulk load of the recorded operand Z of local tile it
            const uint8_t* src = a.op_in + (size_t)((int)blockIdx.x + it * (int)gridDim.x) * op_bytes;
            uint64_t* bo = barO + (it & 1);
            mbar_expect_tx(bo, op_bytes);
            bulk_load(sZ2 + (uint32_t)(it & 1) * bg.a1_bytes, src, op_bytes, bo);
        };
        // ---- P1 of local tile it: gn = g + 0.25 gc (+ tap), Gy = fire * gn -> bf16, residual red.add, zeroing of the consumed tile ----
        auto p1 = [&](int it) {
            int b, y0, x0;
            t2_tile_decode(a.tl, (int)blockIdx.x + it * (int)gridDim.x, b, y0, x0);
            const int gy = y0 + py, gx = x0 + px;
            const bool inimg = gy < H && gx < W;
            const int par = it & 1;
            mbar_wait(barT, (uint32_t)(it & 1));
            if (!a.fm.supplied) mbar_wait(barF + (it & 3), (uint32_t)((it >> 2) & 1));      // fire table of this tile published
            const float fire = a.fm.supplied ? (inimg ? a.fm.supplied[((size_t)b * H + gy) * W + gx] : 0.0f) : sFire4[(it & 3) * 128 + r];
            float gn[8];
            const float* gp = sGn + (8 * eh * T2_TH + py) * T2_TW + px;
            const float* gcp = sGcn + (8 * eh * 4 + (py >> 1)) * 8 + (px >> 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v = 0.0f;
                if (8 * eh + i < C) {
                    v = gp[i * T2_TH * T2_TW];
                    if (NS == 2) v = fmaf(0.25f, gcp[i * 32], v);
                }
                gn[i] = v;
            }
            if (a.g_tap != nullptr && inimg) {            // rgb tap at states[t+1]: rare steps only
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (8 * eh + i < a.tap_c)
                        gn[i] = fmaf(a.tap_scale, __ldg(a.g_tap + ((size_t)b * a.tap_c + 8 * eh + i) * plane + (size_t)gy * W + gx), gn[i]);
            }
            float gyv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { gyv[i] = fire * gn[i]; b2acc[i] += gyv[i]; }
            uint4 pk;
            pk.x = pack_bf16(gyv[0], gyv[1]); pk.y = pack_bf16(gyv[2], gyv[3]); pk.z = pack_bf16(gyv[4], gyv[5]); pk.w = pack_bf16(gyv[6], gyv[7]);
            *reinterpret_cast<uint4*>(sGy2 + (uint32_t)par * 4096u + (uint32_t)eh * 2048u + row_off) = pk;
            // residual path: dL/dx_t += dL/dx_{t+1} (fine + 0.25 coarse + tap), four cells of a channel row per reduction; and the
            // consumed tile of g_{t+1} is zeroed in global memory (it is the output of the next launch)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int i = tid + q * T3_NE;
                const int c = i >> 5, rr = (i >> 2) & 7, x4 = (i & 3) * 4;
                if (c < C && y0 + rr < H && x0 + x4 < W) {
                    float4 v = *reinterpret_cast<const float4*>(sGn + (c * T2_TH + rr) * T2_TW + x4);
                    if (NS == 2) {
                        const float2 gc2 = *reinterpret_cast<const float2*>(sGcn + (c * 4 + (rr >> 1)) * 8 + (x4 >> 1));
                        v.x = fmaf(0.25f, gc2.x, v.x); v.y = fmaf(0.25f, gc2.x, v.y); v.z = fmaf(0.25f, gc2.y, v.z); v.w = fmaf(0.25f, gc2.y, v.w);
                    }
                    const size_t off = ((size_t)b * C + c) * plane + (size_t)(y0 + rr) * W + x0 + x4;
                    if (a.g_tap != nullptr && c < a.tap_c) {
                        const float4 tp = __ldg(reinterpret_cast<const float4*>(a.g_tap + ((size_t)b * a.tap_c + c) * plane + (size_t)(y0 + rr) * W + x0 + x4));
                        v.x = fmaf(a.tap_scale, tp.x, v.x); v.y = fmaf(a.tap_scale, tp.y, v.y); v.z = fmaf(a.tap_scale, tp.z, v.z); v.w = fmaf(a.tap_scale, tp.w, v.w);
                    }
                    red_add_v4(a.g_out + off, v.x, v.y, v.z, v.w);
                    if (a.zero_in) *reinterpret_cast<float4*>(a.g_in + off) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (NS == 2 && a.zero_cin && tid < 128) {
                const int c = tid >> 3, rr = (tid >> 1) & 3, x4 = (tid & 1) * 4;
                if (c < C && (y0 >> 1) + rr < (H >> 1) && (x0 >> 1) + x4 < (W >> 1))
                    *reinterpret_cast<float4*>(a.gc_in + ((size_t)b * C + c) * (plane >> 2) + (size_t)((y0 >> 1) + rr) * (W >> 1) + (x0 >> 1) + x4) =
                        make_float4(0.f, 0.f, 0.f, 0.f);
            }
            fence_proxy_async();          // operand images -> MMA; stage reads before the next TMA write
            mbar_arrive(barG + par);
            if (tid == 0 && it + 1 < n_my) {      // the stage is free once every E thread has read it
                mbar_wait(barG + par, (uint32_t)((it >> 1) & 1));
                issue_g(it + 1);
            }
        };

        if (tid == 0) {
            issue_g(0);
            issue_z(0);
            if (n_my > 1) issue_z(1);
        }
        p1(0);
        for (int it = 0; it < n_my; ++it) {
            const uint32_t R = T3_R0 + 128u * (uint32_t)(it & 1);
            if (tid == 0) T3_STAMP(it, 0);
            if (it >= 1) {
                mbar_wait(barM3b, (uint32_t)((it - 1) & 1));      // H | Ga, Gy / Z of the other buffer are free
                if (tid == 0 && it + 1 < n_my) issue_z(it + 1);
            }
            mbar_wait(barM2, (uint32_t)(it & 1));
            tc_fence_after();
            if (tid == 0) T3_STAMP(it, 1);
            // ---- E1: h = relu(D1), g_a = D3 * [D1 > 0] -> bf16 operands; thread -> hidden units 64 eh .. 64 eh + 63 ----
#pragma unroll
            for (int hp = 0; hp < 4; ++hp) {
                const int j0 = 64 * eh + 16 * hp;
                if (j0 < fc) {
                    uint32_t av[16], gv[16];
                    tmem_ld16(tmem_lane + T3_D1 + (uint32_t)j0, av);
                    tmem_ld16(tmem_lane + R + (uint32_t)j0, gv);
                    tmem_ld_wait();
#pragma unroll
                    for (int qq = 0; qq < 2; ++qq) {
                        uint4 o, p;
                        o.x = pack_bf16_relu(__uint_as_float(av[qq * 8 + 0]), __uint_as_float(av[qq * 8 + 1]));
                        o.y = pack_bf16_relu(__uint_as_float(av[qq * 8 + 2]), __uint_as_float(av[qq * 8 + 3]));
                        o.z = pack_bf16_relu(__uint_as_float(av[qq * 8 + 4]), __uint_as_float(av[qq * 8 + 5]));
                        o.w = pack_bf16_relu(__uint_as_float(av[qq * 8 + 6]), __uint_as_float(av[qq * 8 + 7]));
                        p.x = pack_bf16(__uint_as_float(gv[qq * 8 + 0]), __uint_as_float(gv[qq * 8 + 1])) & bf16x2_nz_mask(o.x);
                        p.y = pack_bf16(__uint_as_float(gv[qq * 8 + 2]), __uint_as_float(gv[qq * 8 + 3])) & bf16x2_nz_mask(o.y);
                        p.z = pack_bf16(__uint_as_float(gv[qq * 8 + 4]), __uint_as_float(gv[qq * 8 + 5])) & bf16x2_nz_mask(o.z);
                        p.w = pack_bf16(__uint_as_float(gv[qq * 8 + 6]), __uint_as_float(gv[qq * 8 + 7])) & bf16x2_nz_mask(o.w);
                        *reinterpret_cast<uint4*>(sH + (uint32_t)(j0 / 8 + qq) * 2048u + row_off) = o;
                        *reinterpret_cast<uint4*>(sGa + (uint32_t)(j0 / 8 + qq) * 2048u + row_off) = p;
                    }
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(barC);
            if (tid == 0) T3_STAMP(it, 2);
            if (it + 1 < n_my) p1(it + 1);
            if (tid == 0) T3_STAMP(it, 3);
        }
        // ---- flush: D4 [fc x 16] -> gW2p[j][c];  D5 [fc x K1] -> gW1p[k][j] (k' -> reference k, perception columns x s0) ----
        mbar_wait(barM3b, (uint32_t)((n_my - 1) & 1));
        tc_fence_after();
        {
            const int j = r;                       // hidden unit = TMEM lane
            uint32_t v[32];
            if (eh == 0) {
                tmem_ld16(tmem_lane + T3_D4, v);
                tmem_ld_wait();
                if (j < fc)
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        if (c < C) atomicAdd(a.gW2p + j * g.CP + c, __uint_as_float(v[c]));
            }
#pragma unroll 1
            for (int k0 = 48 * eh; k0 < 48 * eh + 48; k0 += 16) {      // eh 0: k' 0..47, eh 1: 48..95 (K1 <= 80)
                if (k0 < bg.K1) {
                    tmem_ld16(tmem_lane + T3_D5 + (uint32_t)k0, v);
                    tmem_ld_wait();
                    if (j < fc) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int kp = k0 + i;
                            const int kc = kp >> 3, s = kp & 7;
                            int k = -1;
                            float sc = 1.0f;
                            if (kc < bg.npairs) { const int c = 2 * kc + (s >> 2); if (c < C) { k = (s & 3) * C + c; sc = g.s0; } }
                            else if (kc == bg.npairs) { const int src = dynca_cond_slot_src(g.cc, s); if (src >= 0) k = 4 * C + src; else if (src == -2) k = g.P; }
                            if (k >= 0) atomicAdd(a.gW1p + k * g.FCpad + j, sc * __uint_as_float(v[i]));
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float s = b2acc[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0 && 8 * eh + i < C) atomicAdd(a.gb2p + 8 * eh + i, s);
            }
        }
        tc_fence_before();
    } else {
        // =========================== S warps ===========================
        const int grp = (warp - 8) >> 2, q = warp & 3;           // tile parity handled, lane quarter
        const int P = q >> 1;                                    // tile rows 4P .. 4P+3; coarse footprint rows 3P .. 3P+2
        const int m = (q & 1) * 32 + lane;                       // k' = 4c + f
        const int f = m & 3, c = m >> 2;
        const T3Coef K = t3_coef(f);
        const float2 v0 = make_float2(K.v0, K.v0), v1 = make_float2(K.v1, K.v1), v2 = make_float2(K.v2, K.v2);
        const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        T3Emit e;
        e.H = H; e.W = W; e.pad = g.pad;
        e.b0 = (f & 1) != 0; e.b1 = (f & 2) != 0; e.act = c < C;
        const int Hc = H >> 1, Wc = W >> 1;
        for (int it = grp; it < n_my; it += 2) {
            int b, y0, x0;
            t2_tile_decode(a.tl, (int)blockIdx.x + it * (int)gridDim.x, b, y0, x0);
            const uint32_t R = tmem_lane + T3_R0 + 128u * (uint32_t)grp;
            const uint32_t ph = (uint32_t)((it >> 1) & 1);
            e.y0 = y0; e.x0 = x0;
            e.gob = a.g_out + ((size_t)b * C + (e.act ? c : 0)) * plane;
            e.fast = y0 > 0 && x0 > 0 && y0 + T2_TH < H && x0 + T2_TW < W;
            if (q == 0 && lane == 0) T3_STAMP(it, 8);
            mbar_wait(barM3a + grp, ph);
            tc_fence_after();
            if (q == 0 && lane == 0) T3_STAMP(it, 9);
            // ---- fine transposed stencil of tile rows 4P .. 4P+3 -> output rows 4P-1 .. 4P+4.  One copy of the code (the kernel
            //      must stay inside the instruction cache): iteration k streams input row k (none for k = 4, 5) through the two
            //      pending output rows PA (row k) / PB (row k + 1) and emits output row k - 1; the tensor-memory load of row k + 1 is
            //      in flight during the reduction of row k - 1.  Pairs of ring positions (2j, 2j+1) are packed fp32 pairs. ----
            {
                float2 PA[9], PB[9];
                float of[18];
#pragma unroll
                for (int j = 0; j < 9; ++j) { PA[j] = make_float2(0.f, 0.f); PB[j] = make_float2(0.f, 0.f); }
#pragma unroll
                for (int p = 0; p < 18; ++p) of[p] = 0.0f;
#pragma unroll 1
                for (int k = 0; k < 7; ++k) {
                    uint32_t gv[16];
                    if (k < 4) tmem_ld16(R + 16u * (uint32_t)(4 * P + k), gv);      // in flight during the reduction below
                    if (k >= 1) t3_emit_fine(e, 4 * P + k - 2, of);                   // output row produced by the previous iteration
                    float2 o[9];
                    if (k < 4) {
                        tmem_ld_wait16(gv);
                        float G[20];                             // G[i + 2] = cell i: two zero cells either side
                        G[0] = G[1] = G[18] = G[19] = 0.0f;
#pragma unroll
                        for (int i = 0; i < 16; ++i) G[i + 2] = __uint_as_float(gv[i]);
#pragma unroll
                        for (int j = 0; j < 9; ++j) {
                            // T(p) = h0 G(p) + h1 G(p-1) + h2 G(p-2), p = 2j, 2j+1
                            float2 t;
                            t.x = fmaf(K.h2, G[2 * j], fmaf(K.h1, G[2 * j + 1], K.h0 * G[2 * j + 2]));
                            t.y = fmaf(K.h2, G[2 * j + 1], fmaf(K.h1, G[2 * j + 2], K.h0 * G[2 * j + 3]));
                            o[j] = f2fma(v0, t, PA[j]);
                            PA[j] = f2fma(v1, t, PB[j]);
                            PB[j] = make_float2(K.v2 * t.x, K.v2 * t.y);
                        }
                        // centre term of the Laplacian: position p = cell + 1
#pragma unroll
                        for (int j = 0; j < 9; ++j) {
                            PA[j].x = fmaf(K.ctr, G[2 * j + 1], PA[j].x);
                            PA[j].y = fmaf(K.ctr, G[2 * j + 2], PA[j].y);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 9; ++j) { o[j] = PA[j]; PA[j] = PB[j]; PB[j] = make_float2(0.f, 0.f); }
                    }
#pragma unroll
                    for (int j = 0; j < 9; ++j) { of[2 * j] = o[j].x; of[2 * j + 1] = o[j].y; }
                }
            }
            if (q == 0 && lane == 0) T3_STAMP(it, 10);
            if (NS == 2) {
                // ---- bf16(D6^T) of all 128 cells, packed in place into the first 64 columns: the A operand of D7^T
                //      (row y goes to columns 8y .. 8y+7, always behind the columns still to be read) ----
#pragma unroll 1
                for (int y = 0; y < 8; y += 2) {
                    uint32_t ga[16], gb[16], pk[8];
                    tmem_ld16(R + 16u * (uint32_t)y, ga);
                    tmem_ld16(R + 16u * (uint32_t)y + 16u, gb);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(__uint_as_float(ga[2 * i]), __uint_as_float(ga[2 * i + 1]));
                    tmem_st8(R + 8u * (uint32_t)y, pk);
#pragma unroll
                    for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(__uint_as_float(gb[2 * i]), __uint_as_float(gb[2 * i + 1]));
                    tmem_st8(R + 8u * (uint32_t)y + 8u, pk);
                }
                tmem_st_wait();
            }
            tc_fence_before();
            mbar_arrive(barD + grp);
            if (q == 0 && lane == 0) T3_STAMP(it, 11);
            if (NS == 2) {
                mbar_wait(barM4 + grp, ph);
                tc_fence_after();
                if (q == 0 && lane == 0) T3_STAMP(it, 12);
                // ---- coarse: footprint rows 3P .. 3P+2 (10 columns each) -> output rows 3P-1 .. 3P+3 of the 8 x 12 coarse ring ----
                const int cy0 = (y0 >> 1) - 1, cx0 = (x0 >> 1) - 1;          // coarse coordinates of footprint cell (0, 0)
                float G[3][10];
                if (P == 0) {                                    // columns 0 .. 29 (16-column aligned loads)
                    uint32_t va[16], vb[16];
                    tmem_ld16(R + 64u, va);
                    tmem_ld16(R + 80u, vb);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 30; ++i) G[i / 10][i % 10] = __uint_as_float(i < 16 ? va[i] : vb[i - 16]);
                } else {                                         // columns 30 .. 59
                    uint32_t va[16], vb[16], vc[16];
                    tmem_ld16(R + 80u, va);
                    tmem_ld16(R + 96u, vb);
                    tmem_ld16(R + 112u, vc);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 30; ++i) G[i / 10][i % 10] = __uint_as_float(i < 2 ? va[14 + i] : (i < 18 ? vb[i - 2] : vc[i - 18]));
                }
                const bool cfast = cy0 >= 1 && cx0 >= 1 && cy0 + T2_QH + 1 <= Hc && cx0 + T2_QW + 1 <= Wc;
                if (!cfast) {
                    // transpose of the replicate extension: footprint cells outside the image fold into the clamped cell.
                    // rows: footprint row fr sits at coarse row cy0 + fr; rows below the image fold upwards (chain), row 0 of a top
                    // tile folds into row 1.  Only the rows this pair holds are touched: the single case that would cross the pairs
                    // (a ragged tile with 4 valid rows: row 3 -> row 2) is handled by relocating row 3 below, its rows 4 / 5 are zero.
#pragma unroll
                    for (int k = 2; k >= 1; --k) {
                        if (cy0 + 3 * P + k >= Hc) {
#pragma unroll
                            for (int i = 0; i < 10; ++i) { G[k - 1][i] += G[k][i]; G[k][i] = 0.0f; }
                        }
                    }
                    if (P == 0 && cy0 < 0) {
#pragma unroll
                        for (int i = 0; i < 10; ++i) { G[1][i] += G[0][i]; G[0][i] = 0.0f; }
                    }
#pragma unroll
                    for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
                        for (int i = 9; i >= 1; --i)
                            if (cx0 + i >= Wc) { G[rr][i - 1] += G[rr][i]; G[rr][i] = 0.0f; }
                        if (cx0 < 0) { G[rr][1] += G[rr][0]; G[rr][0] = 0.0f; }
                    }
                }
                T3Emit ec = e;
                ec.fast = cfast;
                float* gcb = a.gc_out + ((size_t)b * C + (e.act ? c : 0)) * (plane >> 2);
                const int rb = (P == 1 && cy0 + 3 >= Hc) ? Hc - 1 : cy0 + 3 * P;      // coarse image row of the pair's first footprint row
                const int oy0 = rb - 1;                          // coarse image row of the first output row
                float2 PA[6], PB[6];
#pragma unroll
                for (int j = 0; j < 6; ++j) { PA[j] = make_float2(0.f, 0.f); PB[j] = make_float2(0.f, 0.f); }
#pragma unroll 1
                for (int k = 0; k < 5; ++k) {
                    float2 o[6];
                    if (k < 3) {
                        float Gk[14];                            // Gk[p + 2]: two zero cells either side
                        Gk[0] = Gk[1] = Gk[12] = Gk[13] = 0.0f;
#pragma unroll
                        for (int i = 0; i < 10; ++i) Gk[i + 2] = k == 0 ? G[0][i] : (k == 1 ? G[1][i] : G[2][i]);
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            float2 t;
                            t.x = fmaf(K.h2, Gk[2 * j], fmaf(K.h1, Gk[2 * j + 1], K.h0 * Gk[2 * j + 2]));
                            t.y = fmaf(K.h2, Gk[2 * j + 1], fmaf(K.h1, Gk[2 * j + 2], K.h0 * Gk[2 * j + 3]));
                            o[j] = f2fma(v0, t, PA[j]);
                            PA[j] = f2fma(v1, t, PB[j]);
                            PB[j] = make_float2(K.v2 * t.x, K.v2 * t.y);
                        }
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            PA[j].x = fmaf(K.ctr, Gk[2 * j + 1], PA[j].x);
                            PA[j].y = fmaf(K.ctr, Gk[2 * j + 2], PA[j].y);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 6; ++j) { o[j] = PA[j]; PA[j] = PB[j]; PB[j] = make_float2(0.f, 0.f); }
                    }
                    float of[12];
#pragma unroll
                    for (int j = 0; j < 6; ++j) { of[2 * j] = o[j].x; of[2 * j + 1] = o[j].y; }
                    t3_emit_coarse(ec, gcb, oy0 + k, cx0 - 1, Hc, Wc, of);
                }
                tc_fence_before();
                mbar_arrive(barE + grp);
                if (q == 0 && lane == 0) T3_STAMP(it, 13);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem_base, 512u);
}

// dL/dx_0 += 0.25 * coarse part (transpose of the 2x2 mean), once per rollout
__global__ void dynca_tc2_add_coarse_kernel(int BC, int H, int W, const float* __restrict__ gc, float* __restrict__ gx) {
    const size_t n = (size_t)BC * H * W;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const size_t bc = i / ((size_t)W * H);
        gx[i] = fmaf(0.25f, __ldg(gc + (bc * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)), gx[i]);
    }
}

// operand images of the BPTT: B2d [N = fc][K = 16] (D3 = Gy . W2), A6 [M = 128][K = fc] = W1h^T with rows 64..127 repeating rows
// 0..63 (row m <-> k' = m % 64 = 4c + filter)
__global__ void dynca_tc3_prep_kernel(DyncaGeom g, const float* __restrict__ w1, const float* __restrict__ w2, __nv_bfloat16* __restrict__ B2d,
                                      __nv_bfloat16* __restrict__ A6) {
    const int n1 = g.fc * 16, n2 = 128 * g.fc;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
        if (i < n1) {
            const int j = i / 16, c = i % 16;
            const float v = c < g.C ? w2[c * g.fc + j] : 0.0f;
            B2d[(size_t)(c >> 3) * (g.fc / 8) * 64 + (size_t)(j >> 3) * 64 + (j & 7) * 8 + (c & 7)] = __float2bfloat16_rn(v);
        } else {
            const int e = i - n1, m = e / g.fc, j = e % g.fc;
            const int kp = m & 63, c = kp >> 2, f = kp & 3;
            const float v = c < g.C ? __bfloat162float(__float2bfloat16_rn(w1[j * g.P + f * g.C + c])) * g.s0 : 0.0f;
            A6[(size_t)(j >> 3) * 1024 + (size_t)m * 8 + (j & 7)] = __float2bfloat16_rn(v);
        }
    }
}

// ---- host side --------------------------------------------------------------------------------------------
bool dynca_tc2_bwd_supported(const DyncaGeom& g) {
    Bf16Geom bg;
    if (!dynca_tc2_supported(g)) return false;
    if (dynca_bf16_geom(g, &bg)) return false;
    return t3_smem(g, bg).total <= 227u * 1024u;
}

// operand images: [forward block of dynca_tc2_prep_weights] then B2d | A6
static inline size_t t3_b2d_bytes(const DyncaGeom& g) { return nca_align_up((size_t)(g.fc / 8) * 256, 256); }
size_t dynca_tc2_bwd_weight_bytes(const DyncaGeom& g) {
    const size_t f = dynca_tc2_weight_bytes(g);
    return f ? f + t3_b2d_bytes(g) + (size_t)(g.fc / 8) * 2048 : 0;
}

int dynca_tc2_prep_bwd_weights(const DyncaGeom& g, const NcaDyncaWeights* w, void* ws, cudaStream_t s) {
    int rc = dynca_tc2_prep_weights(g, w, ws, s);
    if (rc) return rc;
    uint8_t* p = (uint8_t*)ws + dynca_tc2_weight_bytes(g);
    dynca_tc3_prep_kernel<<<32, 256, 0, s>>>(g, w->w1, w->w2, (__nv_bfloat16*)p, (__nv_bfloat16*)(p + t3_b2d_bytes(g)));
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int dynca_tc2_make_gmaps(const DyncaGeom& g, const float* gfine, const float* gcoarse, DyncaTc2Maps* m) {
    int rc = t2_make_map((CUtensorMap*)m->x, gfine, 1, (size_t)g.B * g.C * g.H * g.W, g.B, g.C, g.H, g.W, T2_TH, T2_TW);
    if (rc) return rc;
    if (g.ns == 2) rc = t2_make_map((CUtensorMap*)m->xc, gcoarse, 1, (size_t)g.B * g.C * (g.H / 2) * (g.W / 2), g.B, g.C, g.H / 2, g.W / 2, 4, 8);
    else memcpy(m->xc, m->x, sizeof(m->x));
    return rc;
}

int dynca_tc2_add_coarse(const DyncaGeom& g, const float* gc, float* gx, cudaStream_t s) {
    const size_t n = (size_t)g.B * g.C * g.H * g.W;
    const int grid = (int)((n + 255) / 256 < (size_t)t2_num_sms() * 8 ? (n + 255) / 256 : (size_t)t2_num_sms() * 8);
    dynca_tc2_add_coarse_kernel<<<grid, 256, 0, s>>>(g.B * g.C, g.H, g.W, gc, gx);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int dynca_tc2_backward_step(const DyncaGeom& g, const void* ws, float* wsG, const DyncaTc2Maps* gm, float* g_in, float* gc_in, int zero_in,
                            int zero_cin, const float* g_tap, int tap_c, float tap_scale, float* g_out, float* gc_out, const FireMask& fm,
                            cudaStream_t s, int pdl, const uint8_t* op_in) {
    T3BwdArgs a;
    int rc = dynca_bf16_geom(g, &a.bg);
    if (rc) return rc;
    if (op_in == nullptr) { nca_set_error("the tcgen05 BPTT step needs the recorded perception operand of the step"); return NCA_ERR_ARG; }
    a.g = g; a.op_in = op_in;
    a.g_in = g_in; a.gc_in = gc_in; a.zero_in = zero_in; a.zero_cin = zero_cin;
    a.g_tap = g_tap; a.tap_c = tap_c; a.tap_scale = tap_scale; a.g_out = g_out; a.gc_out = gc_out;
    const uint8_t* wb = (const uint8_t*)ws;
    a.B1 = (const __nv_bfloat16*)wb;
    a.U = (const __nv_bfloat16*)(wb + a.bg.b1_bytes + a.bg.b2_bytes + 64);
    const uint8_t* p = wb + dynca_tc2_weight_bytes(g);
    a.B2d = (const __nv_bfloat16*)p;
    a.A6 = (const __nv_bfloat16*)(p + t3_b2d_bytes(g));
    a.gW1p = wsG; a.gW2p = a.gW1p + (size_t)g.Ppad * g.FCpad; a.gb2p = a.gW2p + (size_t)g.FCpad * g.CP;
    a.fm = fm;
    a.tl = t2_make_tiles(g.B, g.H, g.W);
    static long long* tdbg = nullptr;
    const bool timing = getenv("NCA_T2_TDBG") != nullptr;
    if (timing && !tdbg) { cudaMalloc(&tdbg, 256 * sizeof(long long)); cudaMemset(tdbg, 0, 256 * sizeof(long long)); }
    a.tdbg = timing ? tdbg : nullptr;
    const size_t smem = t3_smem(g, a.bg).total;
    int grid = t2_num_sms();
    if (grid > a.tl.n_tiles) grid = a.tl.n_tiles;
    const CUtensorMap* tg = (const CUtensorMap*)gm->x;
    const CUtensorMap* tgc = (const CUtensorMap*)gm->xc;
#define T3_LAUNCH(NS_, CT_, FT_)                                                                                               \
    do {                                                                                                                        \
        NCA_CUDA_OK(cudaFuncSetAttribute(dynca_bwd_tc3_kernel<NS_, CT_, FT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        NCA_CUDA_OK(t2_launch(dynca_bwd_tc3_kernel<NS_, CT_, FT_>, grid, T3_NTHREADS, smem, s, pdl != 0, *tg, *tgc, a));        \
    } while (0)
#define T3_LAUNCH_CF(CT_, FT_) do { if (g.ns == 2) T3_LAUNCH(2, CT_, FT_); else T3_LAUNCH(1, CT_, FT_); } while (0)
    const bool T2_NOSPEC = t2_nospec("NCA_T2_NOSPEC_BWD");
    T2_DISPATCH_CF(g.C, g.fc, T3_LAUNCH_CF);
    NCA_LAUNCH_OK();
    if (timing) {      // debug only: synchronous dump of CTA 0's phase timestamps (cycles since E's first stamp)
        long long h[256];
        cudaMemcpy(h, tdbg, sizeof(h), cudaMemcpyDeviceToHost);
        const long long t0 = h[0];
        for (int it = 0; it < 8; ++it) {
            fprintf(stderr, "tc3 tile %d  E: top %lld M2 %lld E1done %lld p1done %lld | S: wait %lld D6T %lld fine %lld barD %lld M4 %lld coarse %lld | MMA: barC %lld grad %lld coarse-1 %lld recomp+1 %lld\n", it,
                    h[it * 32 + 0] - t0, h[it * 32 + 1] - t0, h[it * 32 + 2] - t0, h[it * 32 + 3] - t0, h[it * 32 + 8] - t0, h[it * 32 + 9] - t0,
                    h[it * 32 + 10] - t0, h[it * 32 + 11] - t0, h[it * 32 + 12] - t0, h[it * 32 + 13] - t0, h[it * 32 + 16] - t0, h[it * 32 + 17] - t0,
                    h[it * 32 + 18] - t0, h[it * 32 + 19] - t0);
        }
    }
    return NCA_OK;
}
