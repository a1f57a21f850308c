// DyNCA BPTT step on the 5th-gen tensor cores, second generation (8x16 tiles, TMA staging; companion of dynca_tc2.cu).
// Replaces autograd's replay of ExtraChannels/models/dynca.py:113-123 for one step:
//   given the perception operand Z_t the forward recorded (operand history; dynca_tc2.cu: Z = z_fine + up(z_coarse), one bf16
//   operand for both scales) and g = dL/dx_{t+1}  ->  dL/dx_t and the weight gradients.
//
// One CTA per SM = 16 compute warps (512 threads, one 8x16 tile at a time) + 1 MMA / TMA warp.  Row r = py*16 + px of
// every M = 128 operand is TMEM lane r; thread (r, q = tid >> 7) owns cell r and channels 4q .. 4q+3 (= 16 columns of
// the K-permuted perception order k' = 8*(c/2) + 4*(c%2) + filter) or hidden units 32q .. 32q+31.
//
//   TMA   : g_{t+1} tile, coarse part of g_{t+1} (see below); bulk copy of Z (operand history)
//   P1    : g_y = fire * g -> bf16 Gy
//   MMA   : D1 = Z.W1h^T | D3 = Gy.W2                                                                  (recompute, dgrad 2)
//   E1    : h = relu(D1) -> H ; g_a = D3 * [D1 > 0] -> Ga                                             (bf16 operands)
//   MMA   : D4 += H^T.Gy (gW2) | D5 += Ga^T.Z (gW1) | D6 = Ga.W1h (g_z)
//   E2    : D6 -> zero-padded fp32 planes, and -> bf16 D6b (MN-major operand)
//   MMA   : D7 = U^T.D6b (g_z on the coarse footprint: transpose of the x2 bilinear upsample)
//   P5/P6 : transposed perception: fine planes -> red.add into dL/dx_t ; coarse planes -> red.add into a COARSE gradient
//           buffer gc [B,C,H/2,W/2].  The 2x2-mean transpose (0.25 * gc broadcast to the 4 fine cells) is applied when
//           the next BPTT step reads its g tile (and once at the end for dL/dx_0), so the coarse scale costs one
//           reduction per coarse cell instead of four.
//   P7    : the tile of g_{t+1} (and of its coarse part) this CTA consumed is zeroed in place: those buffers are the
//           outputs of the next launch (ping-pong), so no per-step memset is needed.
// Weight gradients stay in TMEM (D4, D5) over all tiles of the CTA and are flushed once with red.add.
// Without an operand history the caller first lets the forward kernel record Z of the step (ops_only, nca_api.cu).
#include "dynca_tc2.cuh"

#define TB_NCOMP 512
#define TB_NTHREADS 544
#define TB_HDR 2048u
// TMEM columns
#define TB_GA 0u      // bf16 Ga (packed pairs, 64 columns): the A operand of D6, read from tensor memory
#define TB_D1 128u
#define TB_D3 256u    // D3, later D6 (256..) and D7 (320..)
#define TB_D4 384u
#define TB_D5 400u
// fp32 plane geometry (zero padded): fine cell (py,px) at [py+1][px+5] of [10][24] (ring position (oy,ox) reads rows
// oy-1..oy+1, columns ox+3..ox+5: the 4 interior outputs ox = 1+4j.. of a thread start at the 16-byte aligned column
// 4j+4; channel planes are 240 floats = 60 16-byte units apart, = 4 mod 8, so the float4 loads of a (4 column groups x
// 2 channels) quarter warp are conflict free); coarse cell (qy,qx) at [qy+2][qx+2] of [10][14]
#define TB_PR 10
#define TB_PS 24
#define TB_PP (TB_PR * TB_PS)
#define TB_CPR 10
#define TB_CPS 14
#define TB_CPP (TB_CPR * TB_CPS)

struct T2BwdArgs {
    DyncaGeom g;
    Bf16Geom bg;
    float* g_in; float* gc_in;                        // dL/dx_{t+1} and its coarse part (read through TMA; zeroed if zero_in)
    int zero_in, zero_cin;
    const float* g_tap; int tap_c; float tap_scale;   // optional rgb tap at states[t+1]
    float* g_out; float* gc_out;                      // dL/dx_t (red.add) and its coarse part (red.add)
    const __nv_bfloat16* B1; const __nv_bfloat16* B2d; const __nv_bfloat16* U;
    float* gW1p; float* gW2p; float* gb2p;            // fp32 accumulators, padded fp32-path layout (red.add)
    FireMask fm;
    T2Tiles tl;
    int pdl;           // launch with the programmatic-serialization attribute (not the first step of a call)
    const uint8_t* op_in;   // operand history of this step (Z per tile, written by the forward)
    long long* tdbg;
};

struct TBSmem {
    uint32_t b1, b2d, u, cpl, gn, gcn, d6b, z, gy, h, ga, px, ctr, cpx, cctr, total;
};
// Z / Gy are double buffered over tiles (the next tile's operands arrive while the gradient MMAs of the current tile run); the
// fine fp32 planes overlay H | Ga (dead once the gradient MMAs are complete).  With two scales the coarse planes have storage of
// their own - their zero pads are written once per launch, and the D7 scatter runs on two otherwise idle warps during the fine
// transposed stencil.
__host__ __device__ static inline TBSmem tb_smem(const DyncaGeom& g, const Bf16Geom& bg) {
    TBSmem s;
    const uint32_t C = (uint32_t)g.C;
    uint32_t o = TB_HDR;
    s.b1 = o; o += bg.b1_bytes;
    s.b2d = o; o += (uint32_t)(g.fc / 8) * 256u;
    s.u = o; o += g.ns == 2 ? 16384u : 0u;
    o = (o + 127u) & ~127u;
    s.cpl = o; o += g.ns == 2 ? 3u * C * TB_CPP * 4u + C * T2_QH * T2_QW * 4u : 0u;      // coarse planes + centre terms
    o = (o + 127u) & ~127u;
    s.gn = o; o += C * T2_TH * T2_TW * 4u;
    o = (o + 127u) & ~127u;
    s.gcn = o; o += g.ns == 2 ? C * 4u * 8u * 4u : 0u;
    o = (o + 127u) & ~127u;
    s.d6b = o; o += g.ns == 2 ? 16u * 1024u : 0u;        // bf16(D6), [k' / 8][128 cells][8]: 8 chunks used, M = 128 views read further
    s.z = o; o += 2u * bg.a1_bytes;
    s.gy = o; o += 2u * 4096u;
    o = (o + 127u) & ~127u;
    s.h = o; o += 16u * 2048u;
    s.ga = o; o += 16u * 2048u;
    uint32_t p = s.h;
    s.px = p; p += 3u * C * TB_PP * 4u;
    s.ctr = p; p += C * T2_TH * T2_TW * 4u;
    s.cpx = s.cpl; s.cctr = s.cpl + 3u * C * TB_CPP * 4u;
    s.total = (o > p ? o : p) + 1024u;                   // + slack for the aliasing reads of the last buffer
    return s;
}

// transposed stencils at ring position (oy, ox) of zero-padded planes X, Y, L (row stride S, cell (py,px) at [py+2][px+2]):
//   sum_{a,b} sobel_x[a][b] X[cell(oy-a, ox-b)] + sobel_y[a][b] Y[..] + (lap[a][b] + 16 delta) L[..]
// for NR vertically adjacent outputs oy0 .. oy0+NR-1 (cells addressed relative to ring coordinates: cell row = oy - a).
template <int NR, int S>
__device__ __forceinline__ void tb_stencil_t(const float* __restrict__ X, const float* __restrict__ Y, const float* __restrict__ L,
                                             int oy0, int ox, float out[NR]) {
    float hx[NR + 2], hy[NR + 2];
#pragma unroll
    for (int k = 0; k < NR + 2; ++k) {
        // input cell row py = oy0 - 2 + k  -> plane row py + 2 = oy0 + k ; columns px = ox - b, b = 2 - j -> ox + j
        const int o = (oy0 + k) * S + ox;
        const float x0 = X[o], x2 = X[o + 2];
        const float l0 = L[o], l1 = L[o + 1], l2 = L[o + 2];
        const float y0 = Y[o], y1 = Y[o + 1], y2 = Y[o + 2];
        hx[k] = (x0 - x2) + fmaf(2.0f, l1, l0 + l2);
        hy[k] = fmaf(2.0f, y1, y0 + y2);
    }
#pragma unroll
    for (int k = 0; k < NR; ++k)      // output oy0+k reads input rows (oy0+k) - a, a = 0,1,2 -> hx[k+2], hx[k+1], hx[k]
        out[k] = fmaf(2.0f, hx[k + 1], hx[k] + hx[k + 2]) + (hy[k] - hy[k + 2]);
}

// horizontal pass of the transposed stencils over one plane row: p = X plane at the thread's aligned column 4j+4, the Y and
// L planes follow at +ps, +2ps.  hx / hy = the 4 interior outputs of the group, ex / ey = the ring column (0 for j = 0,
// 17 for j = 3), which sees only cell column 0 (x[1]) / 15 (x[4]).
__device__ __forceinline__ void tb_hrow(const float* __restrict__ p, int ps, int j, float2 hx[2], float2 hy[2], float& ex, float& ey) {
    const float4 xa = *reinterpret_cast<const float4*>(p);
    const float2 xb = *reinterpret_cast<const float2*>(p + 4);
    const float4 ya = *reinterpret_cast<const float4*>(p + ps);
    const float2 yb = *reinterpret_cast<const float2*>(p + ps + 4);
    const float4 la = *reinterpret_cast<const float4*>(p + 2 * ps);
    const float2 lb = *reinterpret_cast<const float2*>(p + 2 * ps + 4);
    // outputs i = 0..3 as the pairs (0,1), (2,3): hx[i] = (x[i] - x[i+2]) + (2 l[i+1] + (l[i] + l[i+2])), hy[i] = 2 y[i+1] + (y[i] + y[i+2])
    const float2 x01 = make_float2(xa.x, xa.y), x23 = make_float2(xa.z, xa.w);
    const float2 l01 = make_float2(la.x, la.y), l23 = make_float2(la.z, la.w);
    const float2 y01 = make_float2(ya.x, ya.y), y23 = make_float2(ya.z, ya.w);
    hx[0] = f2add(f2sub(x01, x23), f2fma2(make_float2(la.y, la.z), f2add(l01, l23)));
    hx[1] = f2add(f2sub(x23, xb), f2fma2(make_float2(la.w, lb.x), f2add(l23, lb)));
    hy[0] = f2fma2(make_float2(ya.y, ya.z), f2add(y01, y23));
    hy[1] = f2fma2(make_float2(ya.w, yb.x), f2add(y23, yb));
    ex = j == 0 ? la.y - xa.y : xb.x + lb.x;
    ey = j == 0 ? ya.y : yb.x;
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {      // p 16-byte aligned
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {                        // p 8-byte aligned
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// image index that padded coordinate r of an axis of n cells folds back to under the TRANSPOSED padding (-1 = dropped)
__device__ __forceinline__ int tb_fold(int r, int n, int mode) {
    if (r >= 0 && r < n) return r;
    if (mode == NCA_PAD_CONSTANT || r < -1 || r > n) return -1;      // only the 1-cell ring of the image carries gradient
    if (mode == NCA_PAD_CIRCULAR) return r < 0 ? n - 1 : 0;
    if (mode == NCA_PAD_REPLICATE) return r < 0 ? 0 : n - 1;
    return r < 0 ? 1 : n - 2;                                         // reflect
}

template <int NS, int CT, int FT>
__global__ void __launch_bounds__(TB_NTHREADS, 1) dynca_bwd_tc2_kernel(const __grid_constant__ CUtensorMap tm_g,
                                                                        const __grid_constant__ CUtensorMap tm_gc, const T2BwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    DyncaGeom g = a.g;
    Bf16Geom bg = a.bg;
    t2_specialize<CT, FT>(g, bg);
    const TBSmem L = tb_smem(g, bg);
#ifdef NCA_T2_TIMING
    if (a.tdbg && blockIdx.x == 0 && threadIdx.x == 0) a.tdbg[128] = clock64();
#endif
    // MMA-completion barriers (tcgen05.commit), one per batch of a tile, so that the MMA warp may run ahead into the next
    // tile's recompute batch without a barrier ever being two phases ahead of its waiters
    uint64_t* barM2 = reinterpret_cast<uint64_t*>(smem + 8);      // D1, D3
    uint64_t* barM3 = reinterpret_cast<uint64_t*>(smem + 16);     // D4, D5, D6
    uint64_t* barM4 = reinterpret_cast<uint64_t*>(smem + 24);     // D7
    uint64_t* barT = reinterpret_cast<uint64_t*>(smem + 32);      // TMA
    // compute -> MMA warp hand-offs (512 arrivals each)
    uint64_t* barG = reinterpret_cast<uint64_t*>(smem + 48);      // Gy written, stage consumed
    uint64_t* barC = reinterpret_cast<uint64_t*>(smem + 64);      // H, Ga written; D1 / D3 consumed
    uint64_t* barD = reinterpret_cast<uint64_t*>(smem + 72);      // D6b (bf16) written
    uint64_t* barE = reinterpret_cast<uint64_t*>(smem + 80);      // D6 / D7 read back (their TMEM columns are D3's)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 88);
    uint64_t* barO = reinterpret_cast<uint64_t*>(smem + 96);      // [2]: operand-history tile (Z) of buffer 0 / 1 loaded
    float* sFire2 = reinterpret_cast<float*>(smem + 128);               // 2 x 128 floats
    uint8_t* sB1 = smem + L.b1;
    uint8_t* sB2d = smem + L.b2d;
    uint8_t* sU = smem + L.u;
    float* sGn = reinterpret_cast<float*>(smem + L.gn);
    float* sGcn = reinterpret_cast<float*>(smem + L.gcn);
    uint8_t* sD6b = smem + L.d6b;
    uint8_t* sZ2 = smem + L.z;            // 2 x a1_bytes
    uint8_t* sGy2 = smem + L.gy;          // 2 x 4096
    uint8_t* sH = smem + L.h;
    uint8_t* sGa = smem + L.ga;
    float* sPX = reinterpret_cast<float*>(smem + L.px);       // [3][C][10][24]: X, Y, L planes
    float* sCtr = reinterpret_cast<float*>(smem + L.ctr);     // [C][8][16]: g + g_z(id) - 16 g_z(lap)
    float* sCPX = reinterpret_cast<float*>(smem + L.cpx);     // [3][C][10][14]
    float* sCCtr = reinterpret_cast<float*>(smem + L.cctr);   // [C][6][10]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = g.C, H = g.H, W = g.W, fc = g.fc;
    const size_t plane = (size_t)H * W;
    const int n_tiles = a.tl.n_tiles;
    const int N6 = 16 * ((bg.npairs + 1) / 2);                 // perception columns of g_z, padded to the MMA granularity
    const uint32_t stage_bytes = (uint32_t)C * (T2_TH * T2_TW + (NS == 2 ? 32 : 0)) * 4u;
    const uint32_t op_bytes = dynca_tc2_op_tile_bytes(g);

    griddep_launch();
    // ---- one-time setup (independent of the previous launch's output: may overlap its tail) ----
    for (uint32_t i = tid; i < bg.b1_bytes / 16; i += TB_NTHREADS)
        reinterpret_cast<uint4*>(sB1)[i] = __ldg(reinterpret_cast<const uint4*>(a.B1) + i);
    for (uint32_t i = tid; i < (uint32_t)(fc / 8) * 256u / 16; i += TB_NTHREADS)
        reinterpret_cast<uint4*>(sB2d)[i] = __ldg(reinterpret_cast<const uint4*>(a.B2d) + i);
    if (NS == 2)
        for (uint32_t i = tid; i < 16384u / 16; i += TB_NTHREADS)
            reinterpret_cast<uint4*>(sU)[i] = __ldg(reinterpret_cast<const uint4*>(a.U) + i);
    // everything an MMA may read before the tile loop writes it must be finite: clear the dynamic area once (this also writes the
    // zero pads of the persistent coarse planes)
    for (uint32_t i = L.cpl / 16 + tid; i < L.total / 16; i += TB_NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(barM2, 1); mbar_init(barM3, 1); mbar_init(barM4, 1); mbar_init(barT, 1);
        mbar_init(barG, TB_NCOMP); mbar_init(barC, TB_NCOMP); mbar_init(barD, TB_NCOMP); mbar_init(barE, TB_NCOMP);
        mbar_init(barO, 1); mbar_init(barO + 1, 1);
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
        // =========================== MMA / TMA warp ===========================
        // deliberately from the ARGUMENTS, not from the specialised geometry: with every descriptor of this warp a compile-time
        // constant, nvcc 12.9 generated a two-scale (16, 128) instantiation whose recompute / weight-gradient products were wrong
        // (caught by tests/test_dynca_bf16_gpu.py).  The cost is nil: one warp's scalar arithmetic.
        const uint32_t lbo_b1 = (uint32_t)(a.g.fc / 8) * 128u;
        const uint32_t id_fc = umma_idesc_bf16(128, fc);
        const uint32_t id_w2 = umma_idesc_bf16_mn(128, 16), id_w1 = umma_idesc_bf16_mn(128, bg.K1), id_c = umma_idesc_bf16_mn(128, N6);
        const uint32_t id_gz = umma_idesc_bf16(128, N6) | (1u << 16);
        const uint64_t dB1 = umma_desc(smem_u32(sB1), lbo_b1, 128u);
        const uint64_t dB2d = umma_desc(smem_u32(sB2d), lbo_b1, 128u);
        // MN-major views (cells / hidden units become K): LBO = 128 (K groups), SBO = group stride of MN
        const uint64_t dHt = umma_desc(smem_u32(sH), 128u, 2048u), dGat = umma_desc(smem_u32(sGa), 128u, 2048u);
        const uint64_t dUt = umma_desc(smem_u32(sU), 128u, 2048u);                    // U^T: [M = coarse cell][K = cell]
        const uint64_t dD6bt = umma_desc(smem_u32(sD6b), 128u, 2048u);                // bf16(D6) as [N = k'][K = cell]
        const uint64_t dB1t = umma_desc(smem_u32(sB1), 128u, lbo_b1);                 // B1 as [N = k'][K = hidden]
        // double-buffered operands: descriptor of buffer 1 = descriptor of buffer 0 + (bytes >> 4)
        const uint64_t dZ_0 = umma_desc(smem_u32(sZ2), 2048u, 128u), dZt_0 = umma_desc(smem_u32(sZ2), 128u, 2048u);
        const uint64_t dGy_0 = umma_desc(smem_u32(sGy2), 2048u, 128u), dGyt_0 = umma_desc(smem_u32(sGy2), 128u, 2048u);
        const uint64_t oZ = (uint64_t)(bg.a1_bytes >> 4), oGy = (uint64_t)(4096u >> 4);
        const uint64_t sB1k = (uint64_t)((2u * lbo_b1) >> 4);
        const int k1steps = bg.K1 / 16, kfsteps = fc / 16;
        const CUtensorMap* const ptm_g = &tm_g;
        const CUtensorMap* const ptm_gc = &tm_gc;
        uint32_t phC = 0, phD = 0, phG = 0, phE = 0;
        const bool leader = elect_one();
        bool first = true;
#define TB_ISSUE_TMA(tile_)                                                                                              \
    do {                                                                                                                 \
        const int tt_ = (tile_);                                                                                         \
        int tb_, ty_, tx_;                                                                                               \
        t2_tile_decode(a.tl, tt_, tb_, ty_, tx_);                                                                        \
        mbar_expect_tx(barT, stage_bytes);                                                                               \
        tma_load_5d(sGn, ptm_g, barT, tx_, ty_, 0, tb_, 0);                                                              \
        if (NS == 2) tma_load_5d(sGcn, ptm_gc, barT, tx_ >> 1, ty_ >> 1, 0, tb_, 0);                                     \
    } while (0)
        // operand-history tile of sequence index par_ (buffer par_ & 1), one mbarrier per buffer
#define TB_ISSUE_OP(tile_, par_)                                                                                         \
    do {                                                                                                                 \
        const uint8_t* src_ = a.op_in + (size_t)(tile_) * op_bytes;                                                      \
        uint64_t* bo_ = barO + ((par_) & 1);                                                                             \
        mbar_expect_tx(bo_, op_bytes);                                                                                   \
        bulk_load(sZ2 + (uint32_t)((par_) & 1) * bg.a1_bytes, src_, op_bytes, bo_);                                      \
    } while (0)
        if (leader && (int)blockIdx.x < n_tiles) {
            TB_ISSUE_TMA(blockIdx.x);
            TB_ISSUE_OP(blockIdx.x, 0);
            if ((int)(blockIdx.x + gridDim.x) < n_tiles) TB_ISSUE_OP(blockIdx.x + gridDim.x, 1);
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint64_t par = (uint64_t)(it & 1);
            if (it >= 1) {
                // the operand buffer of the previous tile is free once its last reader (the gradient batch) is complete: load the
                // tile after this one into it
                mbar_wait(barM3, (uint32_t)((it - 1) & 1));
                if (leader && tile + (int)gridDim.x < n_tiles) TB_ISSUE_OP(tile + gridDim.x, it + 1);
            }
            const uint64_t dZ = dZ_0 + par * oZ, dZt = dZt_0 + par * oZ;
            const uint64_t dGy = dGy_0 + par * oGy, dGyt = dGyt_0 + par * oGy;
            // ---- recompute batch ----
            mbar_wait(barO + (it & 1), (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            if (leader) {
#pragma unroll 5
                for (int ks = 0; ks < k1steps; ++ks)
                    umma_ss(tmem_base + TB_D1, dZ + (uint64_t)(ks * (4096 >> 4)), dB1 + (uint64_t)ks * sB1k, id_fc, ks > 0);
            }
            mbar_wait(barG, phG);                              // Gy written, stage consumed
            phG ^= 1u;
            if (leader && tile + (int)gridDim.x < n_tiles) TB_ISSUE_TMA(tile + gridDim.x);
            if (it > 0) {                                      // D3's columns held D6 / D7 of the previous tile
                mbar_wait(barE, phE);
                phE ^= 1u;
            }
            tc_fence_after();
            if (leader) {
                umma_ss(tmem_base + TB_D3, dGy, dB2d, id_fc, false);
                umma_commit(barM2);
            }
            mbar_wait(barC, phC);                              // H, Ga written; D1 / D3 consumed
            phC ^= 1u;
            tc_fence_after();
            if (leader) {
                // g_z first (the compute warps wait for it), weight gradients behind it in the same batch
#pragma unroll 8
                for (int ks = 0; ks < kfsteps; ++ks)           // D6 = Ga . W1h  (g_z), Ga from tensor memory
                    umma_ts(tmem_base + TB_D3, tmem_base + TB_GA + 8u * (uint32_t)ks, dB1t + (uint64_t)(ks * (256 >> 4)), id_gz, ks > 0);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {               // 16 cells per instruction
                    const uint64_t o = (uint64_t)(ks * (256 >> 4));
                    umma_ss(tmem_base + TB_D4, dHt + o, dGyt + o, id_w2, !(first && ks == 0));
                    umma_ss(tmem_base + TB_D5, dGat + o, dZt + o, id_w1, !(first && ks == 0));
                }
                umma_commit(barM3);
            }
            first = false;
            if (NS == 2) {
                mbar_wait(barD, phD);                          // D6b (bf16) written
                phD ^= 1u;
                tc_fence_after();
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {           // D7 = U^T . D6b  (g_z on the coarse footprint), 16 cells per instruction
                        const uint64_t o = (uint64_t)(ks * (256 >> 4));
                        umma_ss(tmem_base + TB_D3 + 64u, dUt + o, dD6bt + o, id_c, ks > 0);
                    }
                    umma_commit(barM4);
                }
            }
        }
    } else {
        // =========================== compute warps ===========================
        const int r = tid & 127, qtr = tid >> 7;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t row_off = (uint32_t)r * 16u;
        const int py = r >> 4, px = r & 15;
        uint32_t phM2 = 0, phM3 = 0, phM4 = 0, phT = 0;
        float b2acc[4] = {0.f, 0.f, 0.f, 0.f};
        // zero pads of the fine planes: rows 0 and 9 (columns 4..23 as float4) and columns 4 / 21 of rows 1..8 (a scalar
        // pair): 3*C planes x 18 items over 512 threads = at most 2 items per thread, offsets (in floats) fixed for the launch
        int zoff[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = tid + q * TB_NCOMP;
            int o = -1;
            if (i < 3 * C * 18) {
                const int k = i % 18, pl = i / 18;
                if (k < 10) o = pl * TB_PP + (k < 5 ? 0 : 9) * TB_PS + 4 + 4 * (k % 5);
                else o = (pl * TB_PP + (1 + (k - 10)) * TB_PS + 4) | (1 << 30);      // flag: column pair
            }
            zoff[q] = o;
        }
        // fire decisions (warp 14) of a tile, double buffered over tiles
        auto tables = [&](int tile_, int buf_) {
            if (warp != 14 || a.fm.supplied) return;
            int tb_, ty_, tx_;
            t2_tile_decode(a.tl, tile_, tb_, ty_, tx_);
            t2_fire_tile(a.fm, tb_, ty_, tx_, H, W, lane, sFire2 + buf_ * 128);
        };
        // ---- P1 of a tile: g = dL/dx_{t+1} and Gy (-> barrier G).  Runs for tile i+1 while the gradient MMAs of tile i are in
        //      flight.  itn = iteration index of that tile. ----
        auto p1b = [&](int tile_, int b, int y0, int x0, int itn, float (&gn)[4]) {
            const int gy = y0 + py, gx = x0 + px;
            const bool inimg = gy < H && gx < W;
            const int par = itn & 1;
            uint8_t* sGy = sGy2 + (uint32_t)par * 4096u;
            const float* sFire = sFire2 + par * 128;
            mbar_wait(barT, phT);                             // the gradient stage is this phase's only input
            phT ^= 1u;
#ifdef NCA_T2_TIMING
            if (a.tdbg && blockIdx.x == 0 && tid == 0 && itn >= 1 && itn <= 8) a.tdbg[(itn - 1) * 16 + 14] = clock64();
#endif
            // g of this thread's 4 channels (+ coarse part, + tap); g_y = fire * g -> Gy
            {
                const float fire = a.fm.supplied ? (inimg ? a.fm.supplied[((size_t)b * H + gy) * W + gx] : 0.0f) : sFire[r];
                const int nch = min(4, C - 4 * qtr);          // warp-uniform
                const float* gp = sGn + (4 * qtr * T2_TH + py) * T2_TW + px;
                const float* gcp = sGcn + (4 * qtr * 4 + (py >> 1)) * 8 + (px >> 1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float v = 0.0f;
                    if (i < nch) {
                        v = gp[i * T2_TH * T2_TW];
                        if (NS == 2) v = fmaf(0.25f, gcp[i * 32], v);
                    }
                    gn[i] = v;
                }
                if (a.g_tap != nullptr && inimg) {            // rgb tap at states[t+1]: rare steps only
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (4 * qtr + i < a.tap_c)
                            gn[i] = fmaf(a.tap_scale, __ldg(a.g_tap + ((size_t)b * a.tap_c + 4 * qtr + i) * plane + (size_t)gy * W + gx), gn[i]);
                }
                float gyv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { gyv[i] = fire * gn[i]; b2acc[i] += gyv[i]; }
                uint2 pk;
                pk.x = pack_bf16(gyv[0], gyv[1]); pk.y = pack_bf16(gyv[2], gyv[3]);
                *reinterpret_cast<uint2*>(sGy + (uint32_t)(qtr >> 1) * 2048u + row_off + (uint32_t)(qtr & 1) * 8u) = pk;
            }
            // zero the consumed tile of g_{t+1} in global memory (it is the output of the next launch)
            if (a.zero_in) {
                const int c = tid >> 5, rr = (tid >> 2) & 7, x4 = (tid & 3) * 4;
                if (c < C && y0 + rr < H && x0 + x4 < W)
                    *reinterpret_cast<float4*>(a.g_in + ((size_t)b * C + c) * plane + (size_t)(y0 + rr) * W + x0 + x4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (NS == 2 && a.zero_cin && tid < 128) {
                const int c = tid >> 3, rr = (tid >> 1) & 3, x4 = (tid & 1) * 4;
                if (c < C && (y0 >> 1) + rr < (H >> 1) && (x0 >> 1) + x4 < (W >> 1))
                    *reinterpret_cast<float4*>(a.gc_in + ((size_t)b * C + c) * (plane >> 2) + (size_t)((y0 >> 1) + rr) * (W >> 1) + (x0 >> 1) + x4) =
                        make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (tile_ + (int)gridDim.x < n_tiles) tables(tile_ + gridDim.x, (itn + 1) & 1);
#ifdef NCA_T2_TIMING
            if (a.tdbg && blockIdx.x == 0 && tid == 0 && itn >= 1 && itn <= 8) a.tdbg[(itn - 1) * 16 + 15] = clock64();
#endif
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(barG);
        };

        // ---- P6: transposed coarse perception of one tile (coordinates b_, y0_, x0_) -> red.add into the coarse gradient buffer.
        //      warp = (channel pair, 4-row block), lane = (channel of the pair, column of the 8 x 12 coarse ring):
        //      coarse channel planes are 140 floats = 12 banks apart -> conflict-free reads.  With an operand history the coarse
        //      planes have their own storage, and the phase is deferred into the next tile's wait for its gradient MMAs. ----
        auto p6 = [&](int b_, int y0_, int x0_) {
            const int Hc = H >> 1, Wc = W >> 1;
            const int cy0 = (y0_ >> 1) - 1, cx0 = (x0_ >> 1) - 1;
            const bool border_ = y0_ == 0 || x0_ == 0 || y0_ + T2_TH >= H || x0_ + T2_TW >= W || (y0_ + T2_TH + 4 > H || x0_ + T2_TW + 4 > W);
            const int cp = warp >> 1, vb = warp & 1, hc = lane / 12, ox = lane % 12;
            const int c = 2 * cp + hc;
            if (c < C && lane < 24) {
                const float* X = sCPX + c * TB_CPP;
                const float* Y = X + C * TB_CPP;
                const float* Lp = Y + C * TB_CPP;
                float out[4];
                tb_stencil_t<4, TB_CPS>(X, Y, Lp, 4 * vb, ox, out);
                float* gcb = a.gc_out + ((size_t)b_ * C + c) * (plane >> 2);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int oy = 4 * vb + k;
                    if (oy >= 1 && oy <= T2_QH && ox >= 1 && ox <= T2_QW) out[k] += sCCtr[(c * T2_QH + oy - 1) * T2_QW + ox - 1];
                }
                const int gxc = cx0 - 1 + ox, gyc = cy0 - 1 + 4 * vb;
                if (!border_ || (gxc >= 0 && gxc < Wc && gyc >= 0 && gyc + 4 <= Hc)) {
                    float* p = gcb + (size_t)(cy0 - 1 + 4 * vb) * Wc + cx0 - 1 + ox;
#pragma unroll
                    for (int k = 0; k < 4; ++k) { atomicAdd(p, out[k]); p += Wc; }
                } else {
                    const int tx = tb_fold(cx0 - 1 + ox, Wc, g.pad);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int ty = tb_fold(cy0 - 1 + 4 * vb + k, Hc, g.pad);
                        if (ty >= 0 && tx >= 0) atomicAdd(gcb + (size_t)ty * Wc + tx, out[k]);
                    }
                }
            }
        };

        float gn[4] = {0.f, 0.f, 0.f, 0.f}, gn_next[4] = {0.f, 0.f, 0.f, 0.f};
        int pend_b = -1, pend_y0 = 0, pend_x0 = 0;      // tile whose coarse transposed stencil is still to do (operand history)
        if ((int)blockIdx.x < n_tiles) tables(blockIdx.x, 0);
        bar_sync_n(1, TB_NCOMP);
        int nb = 0, ny0 = 0, nx0 = 0;      // coordinates of the tile whose operands are produced ahead
        if ((int)blockIdx.x < n_tiles) {
            t2_tile_decode(a.tl, blockIdx.x, nb, ny0, nx0);
            p1b(blockIdx.x, nb, ny0, nx0, 0, gn);
        }
#ifdef NCA_T2_TIMING
#define TB_STAMP(k_) do { if (a.tdbg && blockIdx.x == 0 && tid == 0 && iter < 8) a.tdbg[iter * 16 + (k_)] = clock64(); } while (0)
#else
#define TB_STAMP(k_) do { } while (0)
#endif
        int iter = 0;
#ifdef NCA_T2_TIMING
        if (a.tdbg && blockIdx.x == 0 && tid == 0) a.tdbg[129] = clock64();
#endif
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
            const int b = nb, y0 = ny0, x0 = nx0;
            const bool border = y0 == 0 || x0 == 0 || y0 + T2_TH >= H || x0 + T2_TW >= W ||
                                (NS == 2 && (y0 + T2_TH + 4 > H || x0 + T2_TW + 4 > W));
            TB_STAMP(0);
            TB_STAMP(2);
            const int next = tile + (int)gridDim.x;
            if (next < n_tiles) t2_tile_decode(a.tl, next, nb, ny0, nx0);
            mbar_wait(barM2, phM2);
            phM2 ^= 1u;
            tc_fence_after();
            TB_STAMP(3);
            // ---- E1: h = relu(D1), g_a = D3 * [D1 > 0] -> bf16 operands; thread -> hidden units 32q .. 32q+31 ----
            if (32 * qtr < fc) {
                // two passes of 16 columns: D1 and D3 values of a pass are live together (32 registers), so the relu mask
                // is consumed as it is produced instead of being parked in predicate bits
#pragma unroll
                for (int hp = 0; hp < 2; ++hp) {
                    uint32_t av[16], gv[16], ga8[8];
                    tmem_ld16(tmem_lane + TB_D1 + 32u * (uint32_t)qtr + 16u * (uint32_t)hp, av);
                    tmem_ld16(tmem_lane + TB_D3 + 32u * (uint32_t)qtr + 16u * (uint32_t)hp, gv);
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
                        *reinterpret_cast<uint4*>(sH + (uint32_t)(4 * qtr + 2 * hp + qq) * 2048u + row_off) = o;
                        *reinterpret_cast<uint4*>(sGa + (uint32_t)(4 * qtr + 2 * hp + qq) * 2048u + row_off) = p;
                        ga8[4 * qq + 0] = p.x; ga8[4 * qq + 1] = p.y; ga8[4 * qq + 2] = p.z; ga8[4 * qq + 3] = p.w;
                    }
                    // the same 16 hidden units as the A operand of D6 = Ga . W1h in tensor memory (no shared-memory fetch for it)
                    tmem_st8(tmem_lane + TB_GA + 16u * (uint32_t)qtr + 8u * (uint32_t)hp, ga8);
                }
            }
            tmem_st_wait();
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(barC);
            TB_STAMP(4);
            // ---- software pipeline, part 2: rest of the NEXT tile's operands while the gradient MMAs of this one run ----
            if (next < n_tiles) p1b(next, nb, ny0, nx0, iter + 1, gn_next);
            TB_STAMP(5);
            if (NS == 2 && pend_b >= 0) { p6(pend_b, pend_y0, pend_x0); pend_b = -1; }      // previous tile's coarse stencil, under the MMAs
            mbar_wait(barM3, phM3);                            // D6, D4, D5 complete; H | Ga are free
            phM3 ^= 1u;
            tc_fence_after();
            TB_STAMP(6);
            // ---- E2: D6 -> fp32 planes (overlay H | Ga) ----
            {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int o = zoff[q];
                    if (o >= 0) {
                        if (o & (1 << 30)) {
                            float* pr = sPX + (o & ~(1 << 30));
                            pr[0] = 0.0f;
                            pr[17] = 0.0f;
                        } else {
                            *reinterpret_cast<float4*>(sPX + o) = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
                if (16 * qtr < N6) {
                    uint32_t v[16];
                    tmem_ld16(tmem_lane + TB_D3 + 16u * (uint32_t)qtr, v);
                    tmem_ld_wait();
                    if (NS == 2) {
                        // bf16(D6) -> D6b [(k' / 8) * 2048 + cell * 16]: the MN-major B operand of D7 = U^T . D6b
                        uint4 o0, o1;
                        o0.x = pack_bf16(__uint_as_float(v[0]), __uint_as_float(v[1])); o0.y = pack_bf16(__uint_as_float(v[2]), __uint_as_float(v[3]));
                        o0.z = pack_bf16(__uint_as_float(v[4]), __uint_as_float(v[5])); o0.w = pack_bf16(__uint_as_float(v[6]), __uint_as_float(v[7]));
                        o1.x = pack_bf16(__uint_as_float(v[8]), __uint_as_float(v[9])); o1.y = pack_bf16(__uint_as_float(v[10]), __uint_as_float(v[11]));
                        o1.z = pack_bf16(__uint_as_float(v[12]), __uint_as_float(v[13])); o1.w = pack_bf16(__uint_as_float(v[14]), __uint_as_float(v[15]));
                        *reinterpret_cast<uint4*>(sD6b + (uint32_t)(2 * qtr) * 2048u + row_off) = o0;
                        *reinterpret_cast<uint4*>(sD6b + (uint32_t)(2 * qtr + 1) * 2048u + row_off) = o1;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c = 4 * qtr + i;
                        if (c < C) {
                            const int o = (c * TB_PR + py + 1) * TB_PS + px + 5;
                            const float lp = __uint_as_float(v[4 * i + 3]);
                            sPX[o] = __uint_as_float(v[4 * i + 1]);
                            sPX[C * TB_PP + o] = __uint_as_float(v[4 * i + 2]);
                            sPX[2 * C * TB_PP + o] = lp;
                            sCtr[(c * T2_TH + py) * T2_TW + px] = fmaf(-16.0f, lp, gn[i] + __uint_as_float(v[4 * i + 0]));
                        }
                    }
                }
                if (NS == 2) {
                    fence_proxy_async();
                    tc_fence_before();
                    mbar_arrive(barD);
                }
                // D6 read: D3's columns are free again as far as this thread is concerned (D7 is read by warps 12 / 13 alone,
                // during the fine transposed stencil)
                if (NS == 1 || (warp != 12 && warp != 13)) { tc_fence_before(); mbar_arrive(barE); }
            }
            TB_STAMP(7);
            bar_sync_n(1, TB_NCOMP);
            TB_STAMP(8);
            // ---- P5: transposed fine perception -> red.add into dL/dx_t.  A thread owns 4 consecutive columns (one
            //      aligned float4 of the image row): warps 0..7 = channel pair x 4 row pairs x 4 column groups (tile rows
            //      1..8 of the ring, plus the ring columns 0 / 17 from the same registers), warps 8..11 = the ring rows 0 / 9.
            //      Interior positions are cells of this tile: unless the tile is ragged they need no folding and go out as
            //      one vector reduction per row. ----
            {
                const bool ragged = y0 + T2_TH > H || x0 + T2_TW > W;
                if (NS == 2 && (warp == 12 || warp == 13)) {
                    // ---- D7 (coarse footprint rows 0..59, all N6 columns) -> the persistent coarse planes, by the two warps of
                    //      lane quarters 0 / 1 that have nothing to do in this phase ----
                    mbar_wait(barM4, phM4);                    // D7 complete
                    phM4 ^= 1u;
                    tc_fence_after();
                    const int qy = r / T2_QW, qx = r % T2_QW;
#pragma unroll 1
                    for (int q4 = 0; 16 * q4 < N6; ++q4) {
                        uint32_t v[16];
                        tmem_ld16(tmem_lane + TB_D3 + 64u + 16u * (uint32_t)q4, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int c = 4 * q4 + i;
                            if (c < C && r < T2_QH * T2_QW) {
                                const int o = (c * TB_CPR + qy + 2) * TB_CPS + qx + 2;
                                const float lp = __uint_as_float(v[4 * i + 3]);
                                sCPX[o] = __uint_as_float(v[4 * i + 1]);
                                sCPX[C * TB_CPP + o] = __uint_as_float(v[4 * i + 2]);
                                sCPX[2 * C * TB_CPP + o] = lp;
                                sCCtr[(c * T2_QH + qy) * T2_QW + qx] = fmaf(-16.0f, lp, __uint_as_float(v[4 * i + 0]));
                            }
                        }
                    }
                    tc_fence_before();
                    mbar_arrive(barE);                         // D7 read: D3's columns are free again
                }
                if (warp < 12) {
                    const int j = lane & 3;
                    int c, row0, side = 0;
                    if (warp < 8) { c = 2 * warp + ((lane >> 2) & 1); row0 = 2 * (lane >> 3); }
                    else { const int t = tid - 256; c = ((t >> 2) & 1) + 2 * (t >> 4); side = (t >> 3) & 1; row0 = side ? 8 : 1; }
                    if (c < C) {
                        const float* X = sPX + c * TB_PP + row0 * TB_PS + 4 * j + 4;
                        float* gob = a.g_out + ((size_t)b * C + c) * plane;
                        const int gx0 = x0 + 4 * j;                     // image column of the group's first output
                        const int gxe = j == 0 ? x0 - 1 : x0 + T2_TW;   // image column of the ring output (j = 0 / 3 only)
                        if (warp < 8) {
                            // output row kk = 0 / 1 reads plane rows kk .. kk+2: stream the 4 rows through the accumulators
                            float2 p0[2], p1[2];          // the two output rows, as column pairs (0,1), (2,3)
                            float e0, e1;
                            {
                                float2 hx[2], hy[2];
                                float ex, ey;
                                tb_hrow(X, C * TB_PP, j, hx, hy, ex, ey);
                                p0[0] = f2add(hx[0], hy[0]); p0[1] = f2add(hx[1], hy[1]);
                                e0 = ex + ey;
                                tb_hrow(X + TB_PS, C * TB_PP, j, hx, hy, ex, ey);
                                p0[0] = f2fma2(hx[0], p0[0]); p0[1] = f2fma2(hx[1], p0[1]);
                                p1[0] = f2add(hx[0], hy[0]); p1[1] = f2add(hx[1], hy[1]);
                                e0 = fmaf(2.0f, ex, e0); e1 = ex + ey;
                                tb_hrow(X + 2 * TB_PS, C * TB_PP, j, hx, hy, ex, ey);
                                p0[0] = f2add(p0[0], f2sub(hx[0], hy[0])); p0[1] = f2add(p0[1], f2sub(hx[1], hy[1]));
                                p1[0] = f2fma2(hx[0], p1[0]); p1[1] = f2fma2(hx[1], p1[1]);
                                e0 += ex - ey; e1 = fmaf(2.0f, ex, e1);
                                tb_hrow(X + 3 * TB_PS, C * TB_PP, j, hx, hy, ex, ey);
                                p1[0] = f2add(p1[0], f2sub(hx[0], hy[0])); p1[1] = f2add(p1[1], f2sub(hx[1], hy[1]));
                                e1 += ex - ey;
                            }
                            const float o0[4] = {p0[0].x, p0[0].y, p0[1].x, p0[1].y}, o1[4] = {p1[0].x, p1[0].y, p1[1].x, p1[1].y};
                            const float* ctr = sCtr + (c * T2_TH + row0) * T2_TW + 4 * j;
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) {
                                const float4 cv = *reinterpret_cast<const float4*>(ctr + kk * T2_TW);
                                float out[4];
                                out[0] = (kk ? o1[0] : o0[0]) + cv.x; out[1] = (kk ? o1[1] : o0[1]) + cv.y;
                                out[2] = (kk ? o1[2] : o0[2]) + cv.z; out[3] = (kk ? o1[3] : o0[3]) + cv.w;
                                const float eo = kk ? e1 : e0;
                                const int gy = y0 + row0 + kk;
                                if (!ragged) {
                                    red_add_v4(gob + (size_t)gy * W + gx0, out[0], out[1], out[2], out[3]);
                                } else {
                                    const int ty = tb_fold(gy, H, g.pad);
#pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        const int tx = tb_fold(gx0 + i, W, g.pad);
                                        if (ty >= 0 && tx >= 0) atomicAdd(gob + (size_t)ty * W + tx, out[i]);
                                    }
                                }
                                if (j == 0 || j == 3) {
                                    if (!ragged && gxe >= 0 && gxe < W) {      // inside the image: no folding
                                        atomicAdd(gob + (size_t)gy * W + gxe, eo);
                                    } else {
                                        const int ty = tb_fold(gy, H, g.pad), tx = tb_fold(gxe, W, g.pad);
                                        if (ty >= 0 && tx >= 0) atomicAdd(gob + (size_t)ty * W + tx, eo);
                                    }
                                }
                            }
                        } else {
                            // ring row 0 sees only cell row 0 (taps a = 2), ring row 9 only cell row 7 (taps a = 0)
                            float2 hx[2], hy[2];
                            float ex, ey, out[4];
                            tb_hrow(X, C * TB_PP, j, hx, hy, ex, ey);
                            {
                                const float2 a0 = side ? f2add(hx[0], hy[0]) : f2sub(hx[0], hy[0]);
                                const float2 a1 = side ? f2add(hx[1], hy[1]) : f2sub(hx[1], hy[1]);
                                out[0] = a0.x; out[1] = a0.y; out[2] = a1.x; out[3] = a1.y;
                            }
                            const float eo = side ? ex + ey : ex - ey;
                            const int gy = side ? y0 + T2_TH : y0 - 1;
                            if (!ragged && gy >= 0 && gy < H) {                // ring row inside the image: only the corners may fold
                                red_add_v4(gob + (size_t)gy * W + gx0, out[0], out[1], out[2], out[3]);
                                if (j == 0 || j == 3) {
                                    const int tx = tb_fold(gxe, W, g.pad);
                                    if (tx >= 0) atomicAdd(gob + (size_t)gy * W + tx, eo);
                                }
                            } else {
                                const int ty = tb_fold(gy, H, g.pad);
                                if (ty >= 0) {
#pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        const int tx = tb_fold(gx0 + i, W, g.pad);
                                        if (tx >= 0) atomicAdd(gob + (size_t)ty * W + tx, out[i]);
                                    }
                                    if (j == 0 || j == 3) {
                                        const int tx = tb_fold(gxe, W, g.pad);
                                        if (tx >= 0) atomicAdd(gob + (size_t)ty * W + tx, eo);
                                    }
                                }
                            }
                        }
                    }
                }
            }
            TB_STAMP(9);
            if (NS == 2) {
                if (warp != 12 && warp != 13) {
                    mbar_wait(barM4, phM4);                    // D7 complete (D6b in shared memory is free)
                    phM4 ^= 1u;
                    tc_fence_after();
                }
                bar_sync_n(1, TB_NCOMP);                       // every P5 read of the fine planes is done; coarse planes: written
                TB_STAMP(10);
                const int Hc = H >> 1, Wc = W >> 1;
                const int cy0 = (y0 >> 1) - 1, cx0 = (x0 >> 1) - 1;          // coarse coordinates of footprint cell (0,0)
                if (border) {
                    // transpose of the replicate extension: footprint cells outside the image fold into the clamped cell
                    // (rows, then columns; one thread per (array, channel, column / row) so nothing races); a pass is skipped
                    // when the footprint does not leave the image in that direction
                    if (cy0 < 0 || cy0 + T2_QH > Hc) {
                        for (int i = tid; i < 4 * C * T2_QW; i += TB_NCOMP) {
                            const int qx = i % T2_QW, c = (i / T2_QW) % C, arr = i / (T2_QW * C);
                            float* base = arr < 3 ? sCPX + (arr * C + c) * TB_CPP + 2 * TB_CPS + qx + 2 : sCCtr + c * T2_QH * T2_QW + qx;
                            const int S = arr < 3 ? TB_CPS : T2_QW;
                            for (int qy = 0; qy < T2_QH; ++qy) {
                                const int Qy = cy0 + qy;
                                if (Qy >= 0 && Qy < Hc) continue;
                                const int ty = min(max(Qy, 0), Hc - 1) - cy0;
                                if (ty >= 0 && ty < T2_QH) base[ty * S] += base[qy * S];
                                base[qy * S] = 0.0f;
                            }
                        }
                        bar_sync_n(1, TB_NCOMP);
                    }
                    if (cx0 < 0 || cx0 + T2_QW > Wc) {
                        for (int i = tid; i < 4 * C * T2_QH; i += TB_NCOMP) {
                            const int qy = i % T2_QH, c = (i / T2_QH) % C, arr = i / (T2_QH * C);
                            float* base = arr < 3 ? sCPX + (arr * C + c) * TB_CPP + (qy + 2) * TB_CPS + 2 : sCCtr + (c * T2_QH + qy) * T2_QW;
                            for (int qx = 0; qx < T2_QW; ++qx) {
                                const int Qx = cx0 + qx;
                                if (Qx >= 0 && Qx < Wc) continue;
                                const int tx = min(max(Qx, 0), Wc - 1) - cx0;
                                if (tx >= 0 && tx < T2_QW) base[tx] += base[qx];
                                base[qx] = 0.0f;
                            }
                        }
                        bar_sync_n(1, TB_NCOMP);
                    }
                }
                TB_STAMP(11);
                pend_b = b; pend_y0 = y0; pend_x0 = x0;        // the coarse transposed stencil is deferred into the next tile's MMA wait
            }
            TB_STAMP(12);
            bar_sync_n(1, TB_NCOMP);     // the planes overlay H | Ga, which the next tile's E1 writes
            TB_STAMP(13);
#pragma unroll
            for (int i = 0; i < 4; ++i) gn[i] = gn_next[i];
        }
        if (NS == 2 && pend_b >= 0) p6(pend_b, pend_y0, pend_x0);      // last tile
#ifdef NCA_T2_TIMING
        if (a.tdbg && blockIdx.x == 0 && tid == 0) { a.tdbg[130] = clock64(); a.tdbg[132] = iter; }
#endif
        // ---- flush: D4 [fc x 16] -> gW2p[j][c];  D5 [fc x K1] -> gW1p[k][j] (k' -> reference k, perception columns x s0) ----
        {
            const int j = (warp & 3) * 32 + lane;
            uint32_t v[32];
            if (qtr == 0) {
                tmem_ld16(tmem_lane + TB_D4, v);
                tmem_ld_wait();
                if (j < fc)
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        if (c < C) atomicAdd(a.gW2p + j * g.CP + c, __uint_as_float(v[c]));
            } else {
                const int k0 = 32 * (qtr - 1);      // qtr 1: 0..31, 2: 32..63, 3: 64..79
                if (k0 < bg.K1) {
                    if (bg.K1 - k0 >= 32) tmem_ld32(tmem_lane + TB_D5 + (uint32_t)k0, v);
                    else tmem_ld16(tmem_lane + TB_D5 + (uint32_t)k0, v);
                    tmem_ld_wait();
                    if (j < fc) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int kp = k0 + i;
                            if (kp >= bg.K1 || (i >= 16 && bg.K1 - k0 < 32)) continue;
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
            for (int i = 0; i < 4; ++i) {
                float s = b2acc[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0 && 4 * qtr + i < C) atomicAdd(a.gb2p + 4 * qtr + i, s);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
#ifdef NCA_T2_TIMING
    if (a.tdbg && blockIdx.x == 0 && threadIdx.x == 0) a.tdbg[131] = clock64();
#endif
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

// B2d [N = fc][K = 16] operand image of the dgrad-2 GEMM (D3 = Gy . W2)
__global__ void dynca_tc2_prep_b2d_kernel(DyncaGeom g, const float* __restrict__ w2, __nv_bfloat16* __restrict__ B2d) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.fc * 16; i += gridDim.x * blockDim.x) {
        const int j = i / 16, c = i % 16;
        const float v = c < g.C ? w2[c * g.fc + j] : 0.0f;
        B2d[(size_t)(c >> 3) * (g.fc / 8) * 64 + (size_t)(j >> 3) * 64 + (j & 7) * 8 + (c & 7)] = __float2bfloat16_rn(v);
    }
}

// ---- host side --------------------------------------------------------------------------------------------
bool dynca_tc2_bwd_supported(const DyncaGeom& g) {
    Bf16Geom bg;
    if (!dynca_tc2_supported(g)) return false;
    if (dynca_bf16_geom(g, &bg)) return false;
    return tb_smem(g, bg).total <= 227u * 1024u;
}

// operand images: [forward block of dynca_tc2_prep_weights (B1 | B2 | b2 | U | I)] then B2d
size_t dynca_tc2_bwd_weight_bytes(const DyncaGeom& g) {
    const size_t f = dynca_tc2_weight_bytes(g);
    return f ? f + nca_align_up((size_t)(g.fc / 8) * 256, 256) : 0;
}

int dynca_tc2_prep_bwd_weights(const DyncaGeom& g, const NcaDyncaWeights* w, void* ws, cudaStream_t s) {
    int rc = dynca_tc2_prep_weights(g, w, ws, s);
    if (rc) return rc;
    dynca_tc2_prep_b2d_kernel<<<8, 256, 0, s>>>(g, w->w2, (__nv_bfloat16*)((uint8_t*)ws + dynca_tc2_weight_bytes(g)));
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
    T2BwdArgs a;
    a.pdl = pdl;
    a.op_in = op_in;
    int rc = dynca_bf16_geom(g, &a.bg);
    if (rc) return rc;
    if (op_in == nullptr) { nca_set_error("the tcgen05 BPTT step needs the recorded perception operand of the step"); return NCA_ERR_ARG; }
    a.g = g;
    a.g_in = g_in; a.gc_in = gc_in; a.zero_in = zero_in; a.zero_cin = zero_cin;
    a.g_tap = g_tap; a.tap_c = tap_c; a.tap_scale = tap_scale; a.g_out = g_out; a.gc_out = gc_out;
    a.B1 = (const __nv_bfloat16*)ws;
    a.U = (const __nv_bfloat16*)((const uint8_t*)ws + a.bg.b1_bytes + a.bg.b2_bytes + 64);
    a.B2d = (const __nv_bfloat16*)((const uint8_t*)ws + dynca_tc2_weight_bytes(g));
    a.gW1p = wsG; a.gW2p = a.gW1p + (size_t)g.Ppad * g.FCpad; a.gb2p = a.gW2p + (size_t)g.FCpad * g.CP;
    a.fm = fm;
    a.tl = t2_make_tiles(g.B, g.H, g.W);
    static long long* tdbg = nullptr;
    const bool timing = getenv("NCA_T2_TDBG") != nullptr;
    if (timing && !tdbg) cudaMalloc(&tdbg, 160 * sizeof(long long));
    a.tdbg = timing ? tdbg : nullptr;
    const size_t smem = tb_smem(g, a.bg).total;
    int grid = t2_num_sms();
    if (grid > a.tl.n_tiles) grid = a.tl.n_tiles;
    const CUtensorMap* tg = (const CUtensorMap*)gm->x;
    const CUtensorMap* tgc = (const CUtensorMap*)gm->xc;
#define TB_LAUNCH(NS_, CT_, FT_)                                                                                              \
    do {                                                                                                                      \
        NCA_CUDA_OK(cudaFuncSetAttribute(dynca_bwd_tc2_kernel<NS_, CT_, FT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        NCA_CUDA_OK(t2_launch(dynca_bwd_tc2_kernel<NS_, CT_, FT_>, grid, TB_NTHREADS, smem, s, a.pdl != 0, *tg, *tgc, a));      \
    } while (0)
    // specialised instantiations for the reference's (C, fc) pairs; everything else is generic
#define TB_LAUNCH_CF(CT_, FT_) do { if (g.ns == 2) TB_LAUNCH(2, CT_, FT_); else TB_LAUNCH(1, CT_, FT_); } while (0)
    const bool T2_NOSPEC = t2_nospec("NCA_T2_NOSPEC_BWD");
    T2_DISPATCH_CF(g.C, g.fc, TB_LAUNCH_CF);
    NCA_LAUNCH_OK();
    if (timing) {      // debug only: synchronous dump of CTA 0's phase timestamps
        long long h[160];
        cudaMemcpy(h, tdbg, sizeof(h), cudaMemcpyDeviceToHost);
        fprintf(stderr, "tc2 bwd CTA 0: setup %lld, tile loop %lld (%lld tiles), flush %lld cycles\n", h[129] - h[128], h[130] - h[129], h[132], h[131] - h[130]);
        for (int it = 0; it < 8; ++it) {
            fprintf(stderr, "tc2 bwd timing iter %d:", it);
            for (int k = 0; k < 16; ++k) fprintf(stderr, " %lld", h[it * 16 + k] - h[it * 16]);
            fprintf(stderr, "  | since prev tile start %lld\n", it ? h[it * 16] - h[(it - 1) * 16] : 0);
        }
    }
    return NCA_OK;
}
