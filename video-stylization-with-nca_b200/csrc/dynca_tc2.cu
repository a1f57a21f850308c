// DyNCA step on the 5th-gen tensor cores, second generation (NCA_PREC_BF16 when the shape allows it):
//   8x16 tiles, TMA staging, register-blocked perception, coarse scale through the tensor cores.
// Reference semantics: ExtraChannels/models/dynca.py:71-128 (same step as dynca_f32.cu / dynca_bf16.cu).
//
// One CTA = 256 threads = one 8x16 tile of cells (row r = py*16 + px of every M=128 operand, TMEM lane r); warps w and
// w+4 share the lane quarter w%4 and split columns.  Per tile:
//   TMA  : state tile + 1-cell ring [C][10][20] fp32 (zero filled outside the image; other padding modes are patched
//          on border tiles) and, for two perception scales, the coarse (2x2-mean) tile + 2-cell ring [C][8][12]
//   fine perception (fp32, 4 vertically adjacent cells x 2 channels per thread, separable Sobel / Laplacian)
//          -> A1 [128 x K1] bf16, K-major UMMA layout (K order k' = 8*(c/2) + 4*(c%2) + filter, then the cond chunk)
//   coarse perception -> Zc [64 x 64] bf16 (60 coarse cells of the 6x10 footprint, replicate-extended at the image
//          border so that the x2 bilinear upsample, dynca.py:93-94, becomes a constant matrix U)
//   MMA  : Dz = U . Zc + A1 . I  (fp32, tensor memory)          -> bf16 -> Z = z_fine + up(z_coarse), in place over A1: the ONE
//          perception operand of the step (perceive_multiscale's sum, dynca.py:99-111; the 1 / n_scales is in W1h = W1 / n_scales),
//          which is also what the operand history records for the BPTT (dynca_tc3_bwd.cu)
//          D1 = Z . W1h^T                                      = W1h (z_fine + up(z_coarse)) + cond / bias terms
//   E1   : h = relu(D1) -> bf16 A2 ;  MMA: D2 = A2 . W2^T ;  E2: x' = x + (D2 + b2) * fire, coalesced NCHW store, and
//          the 2x2 means of x' (the coarse state of the next step) by warp shuffles.
#include <mutex>
#include "dynca_tc2.cuh"

struct T2FwdArgs {
    DyncaGeom g;
    Bf16Geom bg;
    const float* cond;
    const float* x_in;      // states slot of this step (border patches, supplied-mask independent)
    const float* xc_in;     // coarse state of this step (border patches)
    float* x_out;
    float* xc_out;          // coarse state of the next step (NS == 2), may be NULL
    int slot_in, cslot_in;  // slot coordinates of x_in / xc_in in the tensor maps
    const __nv_bfloat16* B1; const __nv_bfloat16* B2; const float* b2p; const __nv_bfloat16* U;
    FireMask fm;
    T2Tiles tl;
    uint8_t* op_out;   // operand history of this step (Z of every tile, for the BPTT) or NULL
    const __nv_bfloat16* Iz;   // identity operand [64][64] (two scales)
    int dbg;
    int pdl;           // launch with the programmatic-serialization attribute (not the first step of a call)
    long long* tdbg;   // optional phase timestamps of CTA 0 (debug)
};

struct T2Smem {
    uint32_t b1, b2w, u, iz, x, xc, cond, uni, a1, zc, total;
};
__host__ __device__ static inline T2Smem t2_smem(const DyncaGeom& g, const Bf16Geom& bg) {
    T2Smem s;
    uint32_t o = T2_HDR;
    s.b1 = o; o += bg.b1_bytes;
    s.b2w = o; o += bg.b2_bytes;
    s.u = o; o += g.ns == 2 ? 16384u : 0u;
    s.iz = o; o += g.ns == 2 ? 8192u : 0u;
    o = (o + 127u) & ~127u;
    s.x = o; o += (uint32_t)g.C * T2_XR * T2_XS * 4u;
    o = (o + 127u) & ~127u;
    s.xc = o; o += g.ns == 2 ? (uint32_t)g.C * T2_CR * T2_CS * 4u : 0u;
    o = (o + 127u) & ~127u;
    s.cond = o; o += g.cond_kind == NCA_COND_TENSOR ? (uint32_t)g.cc * T2_TH * T2_TW * 4u : 0u;
    o = (o + 127u) & ~127u;
    s.uni = o;
    s.a1 = o;
    s.zc = s.a1 + bg.a1_bytes;
    o += bg.a1_bytes + (g.ns == 2 ? 8192u + 1024u : 0u);       // the hidden layer (A2) lives in tensor memory; + slack for the M = 128 reads of Zc
    s.total = o;
    return s;
}

#define T2_NTHREADS 288     // 8 compute warps + 1 MMA / TMA warp

// OPS: record the perception operand and stop (no MLP, no state written): the BPTT's recompute path, generic geometry only
template <int NS, int CT, int FT, bool OPS>
__global__ void __launch_bounds__(T2_NTHREADS, NS == 2 ? 2 : 4) dynca_fwd_tc2_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                     const __grid_constant__ CUtensorMap tm_xc,
                                                                     const __grid_constant__ CUtensorMap tm_c, const T2FwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    DyncaGeom g = a.g;
    Bf16Geom bg = a.bg;
    t2_specialize<CT, FT>(g, bg);
    const T2Smem L = t2_smem(g, bg);
    uint64_t* barM = reinterpret_cast<uint64_t*>(smem);         // MMA batch complete (tcgen05.commit)
    uint64_t* barT = reinterpret_cast<uint64_t*>(smem + 8);     // TMA complete
    uint64_t* barA = reinterpret_cast<uint64_t*>(smem + 16);    // 256 arrivals: A1 / Zc written, stage consumed
    uint64_t* barB = reinterpret_cast<uint64_t*>(smem + 24);    // 256 arrivals: Z written (two scales)
    uint64_t* barC = reinterpret_cast<uint64_t*>(smem + 32);    // 256 arrivals: A2 written
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 40);
    uint64_t* barZ = reinterpret_cast<uint64_t*>(smem + 48);    // 256 arrivals: Zc written (two scales: the coarse MMA starts under the fine perception)
    float* sB2 = reinterpret_cast<float*>(smem + 64);           // 16 floats
    float* sFire2 = reinterpret_cast<float*>(smem + 128);       // 2 x 128 floats (double buffered over tiles)
    uint32_t* sCpe2 = reinterpret_cast<uint32_t*>(smem + 128 + 1024);   // 2 x 24: bf16 hi | lo << 16 of the CPE rows / columns
    uint8_t* sB1 = smem + L.b1;
    uint8_t* sB2w = smem + L.b2w;
    uint8_t* sU = smem + L.u;
    uint8_t* sIz = smem + L.iz;
    float* sX = reinterpret_cast<float*>(smem + L.x);
    float* sXc = reinterpret_cast<float*>(smem + L.xc);
    float* sCond = reinterpret_cast<float*>(smem + L.cond);
    uint8_t* sA1 = smem + L.a1;
    uint8_t* sZc = smem + L.zc;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = g.C, H = g.H, W = g.W, fc = g.fc;
    const size_t plane = (size_t)H * W;
    const uint32_t stage_bytes = (uint32_t)C * T2_XR * T2_XS * 4u + (NS == 2 ? (uint32_t)C * T2_CR * T2_CS * 4u : 0u) +
                                 (g.cond_kind == NCA_COND_TENSOR ? (uint32_t)g.cc * T2_TH * T2_TW * 4u : 0u);
    const uint32_t tmem_cols = NS == 2 ? 256u : 128u;
    const int n_tiles = a.tl.n_tiles;

    griddep_launch();
    // ---- one-time setup (independent of the previous step's output: may overlap its tail) ----
    for (uint32_t i = tid; i < bg.b1_bytes / 16; i += T2_NTHREADS)
        reinterpret_cast<uint4*>(sB1)[i] = __ldg(reinterpret_cast<const uint4*>(a.B1) + i);
    for (uint32_t i = tid; i < bg.b2_bytes / 16; i += T2_NTHREADS)
        reinterpret_cast<uint4*>(sB2w)[i] = __ldg(reinterpret_cast<const uint4*>(a.B2) + i);
    if (NS == 2) {
        for (uint32_t i = tid; i < 16384u / 16; i += T2_NTHREADS)
            reinterpret_cast<uint4*>(sU)[i] = __ldg(reinterpret_cast<const uint4*>(a.U) + i);
        for (uint32_t i = tid; i < 8192u / 16; i += T2_NTHREADS)
            reinterpret_cast<uint4*>(sIz)[i] = __ldg(reinterpret_cast<const uint4*>(a.Iz) + i);
    }
    if (tid < 16) sB2[tid] = a.b2p[tid];
    if (tid == 0) {
        mbar_init(barM, 1);
        mbar_init(barT, 1);
        mbar_init(barA, 256);
        mbar_init(barB, 256);
        mbar_init(barC, 256);
        mbar_init(barZ, 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc(tmem_slot, tmem_cols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();          // the previous step's state (and coarse state) is complete and visible from here on
    // columns, two scales: Dz 0..63, D1 128..255; D2 (0..15) reuses Dz's columns (dead once Z is written), the bf16 hidden layer
    // A2 sits at 64..127 (8 columns per K step).  One scale (128 columns, so that three CTAs fit an SM): A2 is written
    // IN PLACE over D1 - a thread owns 64 columns of its lane, reads 32 of them and then stores the 16 packed columns over
    // the part it has already read: K step ks sits at column 64*(ks/4) + 8*(ks%4); D2 goes to columns 32..47 (read by then).
    const uint32_t TM_D1 = NS == 2 ? 128u : 0u, TM_DC = 0u, TM_D2 = NS == 2 ? 0u : 32u;
    if (warp == 8) {
        // =========================== MMA / TMA warp ===========================
        const uint32_t idesc1 = umma_idesc_bf16(128, fc), idesc2 = umma_idesc_bf16(128, 16);
        const int N6 = 16 * ((bg.npairs + 1) / 2);                          // perception columns, padded to the MMA granularity
        const uint32_t idescZ = umma_idesc_bf16(128, N6), idescUz = idescZ | (1u << 16);      // Dz; B operand (Zc) MN-major
        const uint32_t lbo_b1 = (uint32_t)(fc / 8) * 128u;
        // every operand lives at a fixed shared-memory address: descriptors are kernel constants, K steps are adds
        const uint64_t dA1 = umma_desc(smem_u32(sA1), 2048u, 128u), dB1 = umma_desc(smem_u32(sB1), lbo_b1, 128u);
        const uint64_t dZc = umma_desc(smem_u32(sZc), 1024u, 128u), dU = umma_desc(smem_u32(sU), 2048u, 128u);
        const uint64_t dZcB = umma_desc(smem_u32(sZc), 128u, 1024u);      // Zc as [K = coarse cell][N = k'], MN-major B
        const uint64_t dIz = umma_desc(smem_u32(sIz), 1024u, 128u);
        const uint64_t dB2 = umma_desc(smem_u32(sB2w), 256u, 128u);
        const uint64_t sB1k = (uint64_t)((2u * lbo_b1) >> 4);
        const int k1steps = bg.K1 / 16, kzsteps = N6 / 16, k2steps = fc / 16;
        const uint32_t op_bytes = dynca_tc2_op_tile_bytes(g);
        const CUtensorMap* const ptm_x = &tm_x;
        const CUtensorMap* const ptm_xc = &tm_xc;
        const CUtensorMap* const ptm_c = &tm_c;
        uint32_t phA = 0, phB = 0, phC = 0, phZ = 0;
        const bool leader = elect_one();
#define T2_ISSUE_TMA(tile_)                                                                                              \
    do {                                                                                                                 \
        const int tt_ = (tile_);                                                                                         \
        int tb_, ty_, tx_;                                                                                               \
        t2_tile_decode(a.tl, tt_, tb_, ty_, tx_);                                                                        \
        mbar_expect_tx(barT, stage_bytes);                                                                               \
        tma_load_5d(sX, ptm_x, barT, tx_ - 4, ty_ - 1, 0, tb_, a.slot_in);                                               \
        if (NS == 2) tma_load_5d(sXc, ptm_xc, barT, (tx_ >> 1) - 4, (ty_ >> 1) - 2, 0, tb_, a.cslot_in);                 \
        if (g.cond_kind == NCA_COND_TENSOR) tma_load_5d(sCond, ptm_c, barT, tx_, ty_, 0, tb_, 0);                        \
    } while (0)
        if (leader && (int)blockIdx.x < n_tiles) T2_ISSUE_TMA(blockIdx.x);
#ifdef NCA_T2_TIMING
#define T2_MSTAMP(k_) do { if (a.tdbg && blockIdx.x == 0 && leader && miter < 8) a.tdbg[128 + miter * 8 + (k_)] = clock64(); } while (0)
#else
#define T2_MSTAMP(k_) do { } while (0)
#endif
        int miter = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++miter) {
            T2_MSTAMP(0);
            if (NS == 2) {
                // Dz = U . Zc as soon as the coarse operand is there (the compute warps go on with the fine perception) ...
                mbar_wait(barZ, phZ);
                phZ ^= 1u;
                tc_fence_after();
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_ss(tmem_base + TM_DC, dU + (uint64_t)(ks * (4096 >> 4)), dZcB + (uint64_t)(ks * (256 >> 4)), idescUz, ks > 0);
                }
            }
            mbar_wait(barA, phA);
            phA ^= 1u;
            tc_fence_after();
            T2_MSTAMP(1);
            auto gemm1 = [&]() {      // D1 = (A1 | Z) . W1h^T, then the operand-history copy of the operand (one bulk copy per tile)
                if (!OPS) {
#pragma unroll 5
                    for (int ks = 0; ks < k1steps; ++ks)
                        umma_ss(tmem_base + TM_D1, dA1 + (uint64_t)(ks * (4096 >> 4)), dB1 + (uint64_t)ks * sB1k, idesc1, ks > 0);
                    umma_commit(barM);
                }
                if (a.op_out) bulk_store(a.op_out + (size_t)tile * op_bytes, sA1, op_bytes);
            };
            if (leader) {
                if (NS == 2) {
                    // ... + A1 . I: the fine perception joins the accumulator exactly (bf16 x 1 in fp32)
#pragma unroll 4
                    for (int ks = 0; ks < kzsteps; ++ks)
                        umma_ss(tmem_base + TM_DC, dA1 + (uint64_t)(ks * (4096 >> 4)), dIz + (uint64_t)(ks * (2048 >> 4)), idescZ, true);
                    umma_commit(barM);
                } else {
                    gemm1();
                }
                if (tile + (int)gridDim.x < n_tiles) T2_ISSUE_TMA(tile + gridDim.x);     // the stage is free
            }
            T2_MSTAMP(2);
            if (NS == 2) {
                mbar_wait(barB, phB);                           // Z written over A1
                phB ^= 1u;
                tc_fence_after();
                T2_MSTAMP(3);
                if (leader) gemm1();
            }
            if (leader && OPS) {
                bulk_store_wait_read();
                mbar_arrive(barM);                              // the operand buffer is free (no MMA reads it in this mode)
            }
            if (OPS) continue;
            T2_MSTAMP(4);
            mbar_wait(barC, phC);
            phC ^= 1u;
            tc_fence_after();
            T2_MSTAMP(5);
            if (leader) {
                if (a.op_out) bulk_store_wait_read();             // A1 / Zc have been read: the next tile may overwrite them
#pragma unroll 8
                for (int ks = 0; ks < k2steps; ++ks)
                    umma_ts(tmem_base + TM_D2, tmem_base + (NS == 2 ? 64u + 8u * (uint32_t)ks : 64u * (uint32_t)(ks >> 2) + 8u * (uint32_t)(ks & 3)),
                            dB2 + (uint64_t)(ks * (512 >> 4)), idesc2, ks > 0);
                umma_commit(barM);
            }
            T2_MSTAMP(6);
        }
        // the operand-history copies of this CTA are complete (not only read) before it exits
        if (leader && a.op_out) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
        // =========================== compute warps ===========================
        const int r = tid & 127, half = tid >> 7;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t row_off = (uint32_t)r * 16u;
        uint32_t phM = 0, phT = 0;
        const int py = r >> 4, px = r & 15;
        // CPE values of the 8 rows / 16 columns of a tile (warp 7, lanes 0..23), double buffered over tiles
#define T2_CPE_TABLE(tile_, buf_)                                                                                        \
    do {                                                                                                                 \
        if (g.cond_kind == NCA_COND_CPE && warp == 7 && lane < T2_TH + T2_TW) {                                          \
            const int tt_ = (tile_);                                                                                     \
            int tb_, ty_, tx_;                                                                                           \
            t2_tile_decode(a.tl, tt_, tb_, ty_, tx_);                                                                    \
            const float raw_ = lane < T2_TH ? dynca_cpe(ty_ + lane, H, g.cpe_oh) : dynca_cpe(tx_ + lane - T2_TH, W, g.cpe_ow); \
            const __nv_bfloat16 hi_ = __float2bfloat16_rn(raw_), lo_ = __float2bfloat16_rn(raw_ - __bfloat162float(hi_)); \
            sCpe2[(buf_) * 24 + lane] = (uint32_t)__bfloat16_as_ushort(hi_) | ((uint32_t)__bfloat16_as_ushort(lo_) << 16); \
        }                                                                                                                \
    } while (0)
        if ((int)blockIdx.x < n_tiles) T2_CPE_TABLE(blockIdx.x, 0);
        bar_sync_n(1, 256);
#ifdef NCA_T2_TIMING
#define T2_STAMP(k_) do { if (a.tdbg && blockIdx.x == 0 && tid == 0 && iter < 8) a.tdbg[iter * 16 + (k_)] = clock64(); } while (0)
#else
#define T2_STAMP(k_) do { } while (0)
#endif
        int iter = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
            int b, y0, x0;
            t2_tile_decode(a.tl, tile, b, y0, x0);
            const int gy = y0 + py, gx = x0 + px;
            const bool inimg = gy < H && gx < W;
            // border tile: some staged position (fine ring, or the coarse tile whose ring reaches 4 fine cells further) lies
            // outside the image
            const bool border = y0 == 0 || x0 == 0 || y0 + T2_TH >= H || x0 + T2_TW >= W ||
                                (NS == 2 && (y0 + T2_TH + 4 > H || x0 + T2_TW + 4 > W));
            float* sFire = sFire2 + (iter & 1) * 128;
            const uint32_t* sCpe = sCpe2 + (iter & 1) * 24;
            T2_STAMP(0);
            mbar_wait(barT, phT);
            phT ^= 1u;
            T2_STAMP(1);
            if (border && g.pad != NCA_PAD_CONSTANT) {     // uniform over the compute warps
                t2_patch_border<NS, 256>(g, a.x_in, a.xc_in, b, y0, x0, sX, sXc);
                T2_STAMP(4);
                bar_sync_n(1, 256);
            }
            T2_STAMP(7);
            // ---- fire decisions of the tile ----
            if (a.fm.supplied) {
                if (tid < 128) sFire[r] = inimg ? a.fm.supplied[((size_t)b * H + gy) * W + gx] : 0.0f;
            } else if (warp == (iter & 7)) {
                t2_fire_tile(a.fm, b, y0, x0, H, W, lane, sFire);
            }
            if (NS == 2) {         // coarse operand first: its MMA (Dc) then runs under the fine perception
                t2_coarse_to_zc<8>(g, sXc, sZc, bg.npairs, y0, x0, border, tid, warp, lane);
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(barZ);
            }
            t2_fine_to_a1<8>(sX, sA1, C, bg.npairs, warp, lane);
            T2_STAMP(11);
            // residual state of this thread's cell (channels 8*half ..), cond chunk, zero tail chunks of A1
            float xres[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = 8 * half + i;
                xres[i] = c < C ? sX[(c * T2_XR + py + 1) * T2_XS + T2_XO + px + 1] : 0.0f;
            }
            if (half == 0) {
                uint4 cv;
                if (g.cond_kind == NCA_COND_CPE) {           // [cpe_y hi, cpe_x hi, 1, 1, cpe_y lo, cpe_x lo, 0, 0]
                    const uint32_t ry = sCpe[py], cx = sCpe[T2_TH + px];
                    cv = make_uint4((ry & 0xffffu) | (cx << 16), 0x3F803F80u, (ry >> 16) | (cx & 0xffff0000u), 0u);
                } else {
                    cv = g.cond_kind == NCA_COND_TENSOR ? t2_cond_chunk_smem(g, sCond, r, inimg) : dynca_cond_chunk(g, a.cond, b, gy, gx, inimg);
                }
                *reinterpret_cast<uint4*>(sA1 + (uint32_t)bg.npairs * 2048u + row_off) = cv;
            } else {
                for (int ch = bg.npairs + 1; ch < bg.K1 / 8; ++ch)
                    *reinterpret_cast<uint4*>(sA1 + (uint32_t)ch * 2048u + row_off) = make_uint4(0, 0, 0, 0);
            }
            T2_STAMP(2);
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(barA);                                 // ---- A: operands complete, stage consumed ----
            T2_STAMP(3);
            if (NS == 2) {
                mbar_wait(barM, phM);
                phM ^= 1u;
                tc_fence_after();
                T2_STAMP(5);
                // ---- Dz = z_fine + up(z_coarse) (fp32) -> bf16 -> Z, in place over the perception chunks of A1 (the MMA that read
                //      them is complete); thread -> its row, columns 32 half .. 32 half + 31 ----
                {
                    const int N6 = 16 * ((bg.npairs + 1) / 2);
                    const int j0 = 32 * half;
                    if (j0 < N6) {
                        uint32_t v[32];
                        if (N6 - j0 >= 32) tmem_ld32(tmem_lane + TM_DC + (uint32_t)j0, v);
                        else tmem_ld16(tmem_lane + TM_DC + (uint32_t)j0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            if (j0 + 8 * qq < N6) {
                                uint4 o;
                                o.x = pack_bf16(__uint_as_float(v[qq * 8 + 0]), __uint_as_float(v[qq * 8 + 1]));
                                o.y = pack_bf16(__uint_as_float(v[qq * 8 + 2]), __uint_as_float(v[qq * 8 + 3]));
                                o.z = pack_bf16(__uint_as_float(v[qq * 8 + 4]), __uint_as_float(v[qq * 8 + 5]));
                                o.w = pack_bf16(__uint_as_float(v[qq * 8 + 6]), __uint_as_float(v[qq * 8 + 7]));
                                *reinterpret_cast<uint4*>(sA1 + (uint32_t)(j0 / 8 + qq) * 2048u + row_off) = o;
                            }
                        }
                    }
                }
                T2_STAMP(6);
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(barB);                             // ---- B: Z complete ----
            }
            if (OPS) {                                  // the operand is recorded by the MMA warp; wait until the buffer is free again
                mbar_wait(barM, phM);
                phM ^= 1u;
                continue;
            }
            T2_STAMP(8);
            mbar_wait(barM, phM);
            phM ^= 1u;
            tc_fence_after();
            T2_STAMP(9);
            // ---- E1: relu(D1) -> bf16 -> A2 in tensor memory (the A operand of GEMM 2) ----
#pragma unroll 1
            for (int blk = 0; blk < 2; ++blk) {
                const int j0 = 64 * half + 32 * blk;
                if (j0 < fc) {
                    uint32_t v[32], o[16];
                    tmem_ld32(tmem_lane + TM_D1 + (uint32_t)j0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = pack_bf16_relu(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                    tmem_st16(tmem_lane + (NS == 2 ? 64u + (uint32_t)(j0 >> 1) : 64u * (uint32_t)half + 16u * (uint32_t)blk), o);
                }
            }
            tmem_st_wait();
            T2_STAMP(10);
            if (tile + (int)gridDim.x < n_tiles) T2_CPE_TABLE(tile + gridDim.x, (iter + 1) & 1);   // published by the arrive below
            tc_fence_before();
            mbar_arrive(barC);                                 // ---- C: A2 complete ----
            T2_STAMP(12);
            mbar_wait(barM, phM);
            phM ^= 1u;
            tc_fence_after();
            T2_STAMP(13);
            // ---- E2: x' = x + (D2 + b2) * fire ; coarse state of the next step ----
            {
                uint32_t v[8];
                tmem_ld8(tmem_lane + TM_D2 + 8u * (uint32_t)half, v);
                tmem_ld_wait();
                const float fire = sFire[r];
                float xn[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) xn[i] = fmaf(__uint_as_float(v[i]) + sB2[8 * half + i], fire, xres[i]);
                const int nch = min(8, C - 8 * half);                  // warp-uniform
                {
                    float* xo = a.x_out + ((size_t)b * C + 8 * half) * plane + (size_t)gy * W + gx;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (inimg && i < nch) *xo = xn[i];
                        xo += plane;
                    }
                }
                if (NS == 2 && a.xc_out != nullptr) {
                    const size_t cpl = plane >> 2;
                    float* co = a.xc_out + ((size_t)b * C + 8 * half) * cpl + (size_t)(gy >> 1) * (W >> 1) + (gx >> 1);
                    const bool cst = inimg && (lane & 17) == 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float h2 = xn[i] + __shfl_xor_sync(0xffffffffu, xn[i], 1);          // row pair sums
                        const float s4 = h2 + __shfl_xor_sync(0xffffffffu, h2, 16);
                        if (cst && i < nch) *co = 0.25f * s4;
                        co += cpl;
                    }
                }
            }
            T2_STAMP(14);
            // no CTA barrier here: the next tile's MMAs are gated by barrier A, which every compute thread reaches only
            // after its TMEM reads of this tile; shared operands of this tile were released by the MMA completions
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, tmem_cols);
}

// ---- weight / constant operand packing ------------------------------------------------------------------
// B1 [N = fc][K = K1] with the perception columns scaled by 1/n_scales (exact: power of two), B2, b2, and the
// upsample matrix U [128 cells][64 coarse cells] (K-major A operand)
__global__ void dynca_tc2_prep_kernel(DyncaGeom g, Bf16Geom bg, const float* __restrict__ w1, const float* __restrict__ b1,
                                      const float* __restrict__ w2, const float* __restrict__ b2, __nv_bfloat16* __restrict__ B1,
                                      __nv_bfloat16* __restrict__ B2, float* __restrict__ b2p, __nv_bfloat16* __restrict__ U) {
    const int n1 = bg.K1 * g.fc, n2 = g.fc * 16, n3 = g.ns == 2 ? 128 * 64 : 0, n4 = g.ns == 2 ? 64 * 64 : 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + 16 + n3 + n4; i += gridDim.x * blockDim.x) {
        if (i < n1) {
            const int kp = i / g.fc, j = i % g.fc;
            const int kc = kp >> 3, s = kp & 7;
            float v = 0.0f;
            if (kc < bg.npairs) {
                const int c = 2 * kc + (s >> 2), f = s & 3;
                if (c < g.C) v = __bfloat162float(__float2bfloat16_rn(w1[j * g.P + f * g.C + c])) * g.s0;
            } else if (kc == bg.npairs) {
                const int src = dynca_cond_slot_src(g.cc, s);
                if (src >= 0) v = w1[j * g.P + 4 * g.C + src];
                else if (src == -2) v = __bfloat162float(__float2bfloat16_rn(b1[j]));
                else if (src == -3) v = b1[j] - __bfloat162float(__float2bfloat16_rn(b1[j]));
            }
            B1[(size_t)kc * (g.fc / 8) * 64 + (size_t)(j >> 3) * 64 + (j & 7) * 8 + s] = __float2bfloat16_rn(v);
        } else if (i < n1 + n2) {
            const int e = i - n1, j = e / 16, c = e % 16;
            const float v = c < g.C ? w2[c * g.fc + j] : 0.0f;
            B2[(size_t)(j >> 3) * 128 + (size_t)(c >> 3) * 64 + (c & 7) * 8 + (j & 7)] = __float2bfloat16_rn(v);
        } else if (i < n1 + n2 + 16) {
            const int c = i - n1 - n2;
            b2p[c] = c < g.C ? b2[c] : 0.0f;
        } else if (i >= n1 + n2 + 16 + n3) {
            // identity [N = 64][K = 64], K-major, behind U: A1 . I adds the fine perception to the accumulator of U . Zc
            const int e = i - n1 - n2 - 16 - n3, n = e / 64, k = e % 64;
            U[128 * 64 + (size_t)(k >> 3) * 512 + (size_t)n * 8 + (k & 7)] = __float2bfloat16_rn(n == k ? 1.0f : 0.0f);
        } else {
            const int e = i - n1 - n2 - 16, rr = e / 64, q = e % 64;
            const int py = rr >> 4, px = rr & 15, qy = q / T2_QW, qx = q % T2_QW;
            float wy = 0.0f, wx = 0.0f;
            if (q < T2_QH * T2_QW) {
                const int ly = (py >> 1) + 1, lx = (px >> 1) + 1;
                if (py & 1) wy = qy == ly ? 0.75f : (qy == ly + 1 ? 0.25f : 0.0f);
                else wy = qy == ly ? 0.75f : (qy == ly - 1 ? 0.25f : 0.0f);
                if (px & 1) wx = qx == lx ? 0.75f : (qx == lx + 1 ? 0.25f : 0.0f);
                else wx = qx == lx ? 0.75f : (qx == lx - 1 ? 0.25f : 0.0f);
            }
            U[(size_t)(q >> 3) * 1024 + (size_t)(rr >> 3) * 64 + (rr & 7) * 8 + (q & 7)] = __float2bfloat16_rn(wy * wx);
        }
    }
}

// ---- host side --------------------------------------------------------------------------------------------
// operand history: per step, per tile, the perception operand Z [K1/8 chunks x 128 rows x 16 B]
size_t dynca_tc2_op_hist_bytes(const DyncaGeom& g, int T) {
    const T2Tiles tl = t2_make_tiles(g.B, g.H, g.W);
    return (size_t)T * tl.n_tiles * dynca_tc2_op_tile_bytes(g);
}

bool dynca_tc2_supported(const DyncaGeom& g) {
    Bf16Geom bg;
    if (g.fc % 32 != 0 || g.fc < 32 || g.fc > 128 || g.cc + 2 > 8 || g.C > 16) return false;
    if ((g.W & 3) != 0 || (g.ns == 2 && (g.W & 7) != 0)) return false;      // TMA strides: 16-byte multiples
    if ((long long)g.H * g.W >= (1ll << 30)) return false;
    if (dynca_bf16_geom(g, &bg)) return false;
    return bg.K1 <= 80 && t2_smem(g, bg).total <= 227u * 1024u;
}

size_t dynca_tc2_weight_bytes(const DyncaGeom& g) {
    Bf16Geom bg;
    if (dynca_bf16_geom(g, &bg)) return 0;
    return nca_align_up((size_t)bg.b1_bytes + bg.b2_bytes + 64 + (g.ns == 2 ? 16384 + 8192 : 0), 256);
}

int dynca_tc2_prep_weights(const DyncaGeom& g, const NcaDyncaWeights* w, void* ws, cudaStream_t s) {
    Bf16Geom bg;
    int rc = dynca_bf16_geom(g, &bg);
    if (rc) return rc;
    __nv_bfloat16* B1 = (__nv_bfloat16*)ws;
    __nv_bfloat16* B2 = (__nv_bfloat16*)((uint8_t*)ws + bg.b1_bytes);
    float* b2p = (float*)((uint8_t*)ws + bg.b1_bytes + bg.b2_bytes);
    __nv_bfloat16* U = (__nv_bfloat16*)((uint8_t*)ws + bg.b1_bytes + bg.b2_bytes + 64);
    dynca_tc2_prep_kernel<<<32, 256, 0, s>>>(g, bg, w->w1, w->b1, w->w2, w->b2, B1, B2, b2p, U);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

// tensor maps of one rollout: states [slots][B][C][H][W] and (two scales) coarse states [cslots][B][C][H/2][W/2]
int dynca_tc2_make_maps(const DyncaGeom& g, const float* states, int slots, const float* coarse, int cslots, size_t cslot_floats,
                        const float* cond, DyncaTc2Maps* m) {
    static_assert(sizeof(CUtensorMap) == sizeof(m->x), "tensor map size");
    int rc = t2_make_map((CUtensorMap*)m->x, states, slots, (size_t)g.B * g.C * g.H * g.W, g.B, g.C, g.H, g.W, T2_XR, T2_XS);
    if (rc) return rc;
    if (g.ns == 2) rc = t2_make_map((CUtensorMap*)m->xc, coarse, cslots, cslot_floats, g.B, g.C, g.H / 2, g.W / 2, T2_CR, T2_CS);
    else memcpy(m->xc, m->x, sizeof(m->x));
    if (rc) return rc;
    if (g.cond_kind == NCA_COND_TENSOR)
        rc = t2_make_map((CUtensorMap*)m->cond, cond, 1, (size_t)g.B * g.cc * g.H * g.W, g.B, g.cc, g.H, g.W, T2_TH, T2_TW);
    else memcpy(m->cond, m->x, sizeof(m->x));
    return rc;
}

int dynca_tc2_forward_step(const DyncaGeom& g, const void* ws, const DyncaTc2Maps* m, int slot_in, const float* x_in, float* x_out,
                           int cslot_in, const float* xc_in, float* xc_out, const float* cond, const FireMask& fm, cudaStream_t s, int pdl,
                           uint8_t* op_out, int ops_only) {
    T2FwdArgs a;
    a.pdl = pdl;
    a.op_out = op_out;
    int rc = dynca_bf16_geom(g, &a.bg);
    if (rc) return rc;
    a.g = g; a.cond = cond; a.x_in = x_in; a.xc_in = xc_in; a.x_out = x_out; a.xc_out = xc_out;
    a.slot_in = slot_in; a.cslot_in = cslot_in;
    a.B1 = (const __nv_bfloat16*)ws;
    a.B2 = (const __nv_bfloat16*)((const uint8_t*)ws + a.bg.b1_bytes);
    a.b2p = (const float*)((const uint8_t*)ws + a.bg.b1_bytes + a.bg.b2_bytes);
    a.U = (const __nv_bfloat16*)((const uint8_t*)ws + a.bg.b1_bytes + a.bg.b2_bytes + 64);
    a.Iz = a.U + 128 * 64;
    a.fm = fm;
    { const char* e = getenv("NCA_T2_DBG"); a.dbg = e ? atoi(e) : 0; }
    static long long* tdbg = nullptr;
    const bool timing = getenv("NCA_T2_TDBG") != nullptr;
    if (timing && !tdbg) { cudaMalloc(&tdbg, 256 * sizeof(long long)); }
    a.tdbg = timing ? tdbg : nullptr;
    a.tl = t2_make_tiles(g.B, g.H, g.W);
    const size_t smem = t2_smem(g, a.bg).total;
    const uint32_t tcols = g.ns == 2 ? 256u : 128u;
    // persistent grid = SMs x CTAs that are really co-resident: shared memory, REGISTERS (72 per thread with one scale: three
    // CTAs, not the four that shared memory and TMEM would allow - a fourth wave of CTAs would run alone afterwards) and TMEM.
    // Computed once per (kernel, shared-memory size).
    const CUtensorMap* tx = (const CUtensorMap*)m->x;
    const CUtensorMap* txc = (const CUtensorMap*)m->xc;
    const CUtensorMap* tcn = (const CUtensorMap*)m->cond;
    // (function attributes are per device and the launchers run on several threads: the cache is per device, under a mutex)
    static std::mutex occ_mu;
    static size_t occ_smem_dev[NCA_MAX_DEVICES][10];
    static int occ_val_dev[NCA_MAX_DEVICES][10];
    const bool T2_NOSPEC = t2_nospec("NCA_T2_NOSPEC_FWD");
    const int spec = T2_NOSPEC ? 0 : ((g.C == 16 && g.fc == 128) ? 1 : ((g.C == 12 && g.fc == 96) ? 2 : ((g.C == 13 && g.fc == 96) ? 3 : 0)));
    const int oi = ops_only ? 8 + (g.ns == 2 ? 1 : 0) : (g.ns == 2 ? 4 : 0) + spec;
    const int dev = nca_device_ordinal() % NCA_MAX_DEVICES;
    std::unique_lock<std::mutex> occ_lock(occ_mu);
    size_t* const occ_smem = occ_smem_dev[dev];
    int* const occ_val = occ_val_dev[dev];
    if (occ_val[oi] == 0 || occ_smem[oi] != smem) {
        int o = 0;
#define T2F_ATTR(CT_, FT_, OPS_)                                                                                                    \
    do {                                                                                                                            \
        if (g.ns == 2) {                                                                                                            \
            NCA_CUDA_OK(cudaFuncSetAttribute(dynca_fwd_tc2_kernel<2, CT_, FT_, OPS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            o = t2_occupancy_by_regs(dynca_fwd_tc2_kernel<2, CT_, FT_, OPS_>, T2_NTHREADS);                                         \
        } else {                                                                                                                    \
            NCA_CUDA_OK(cudaFuncSetAttribute(dynca_fwd_tc2_kernel<1, CT_, FT_, OPS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            o = t2_occupancy_by_regs(dynca_fwd_tc2_kernel<1, CT_, FT_, OPS_>, T2_NTHREADS);                                         \
        }                                                                                                                           \
    } while (0)
#define T2F_ATTR_STEP(CT_, FT_) T2F_ATTR(CT_, FT_, false)
        if (ops_only) T2F_ATTR(0, 0, true);
        else T2_DISPATCH_CF(g.C, g.fc, T2F_ATTR_STEP);
        const int by_smem = (int)((227 * 1024) / (smem + 1024));
        if (o > by_smem) o = by_smem;
        occ_val[oi] = o < 1 ? 1 : o;
        occ_smem[oi] = smem;
        if (getenv("NCA_T2_DBG")) fprintf(stderr, "tc2 fwd: ns %d smem %zu -> %d CTAs per SM (registers, shared memory)\n", g.ns, smem, occ_val[oi]);
    }
    int occ = occ_val[oi];
    occ_lock.unlock();
    if (occ > (int)(512u / tcols)) occ = (int)(512u / tcols);
    if (occ < 1) occ = 1;
    int grid = t2_num_sms() * occ;
    if (grid > a.tl.n_tiles) grid = a.tl.n_tiles;
#define T2F_LAUNCH(CT_, FT_, OPS_)                                                                                                  \
    do {                                                                                                                            \
        if (g.ns == 2) NCA_CUDA_OK(t2_launch(dynca_fwd_tc2_kernel<2, CT_, FT_, OPS_>, grid, T2_NTHREADS, smem, s, a.pdl != 0, *tx, *txc, *tcn, a)); \
        else NCA_CUDA_OK(t2_launch(dynca_fwd_tc2_kernel<1, CT_, FT_, OPS_>, grid, T2_NTHREADS, smem, s, a.pdl != 0, *tx, *txc, *tcn, a)); \
    } while (0)
#define T2F_LAUNCH_STEP(CT_, FT_) T2F_LAUNCH(CT_, FT_, false)
    if (ops_only) T2F_LAUNCH(0, 0, true);
    else T2_DISPATCH_CF(g.C, g.fc, T2F_LAUNCH_STEP);
    NCA_LAUNCH_OK();
    if (timing) {      // debug only: synchronous dump of CTA 0's phase timestamps (cycles since the tile's first stamp)
        long long h[256];
        cudaMemcpy(h, tdbg, sizeof(h), cudaMemcpyDeviceToHost);
        for (int it = 0; it < 8; ++it) {
            fprintf(stderr, "tc2 mma-warp iter %d (vs compute tile start):", it);
            for (int k = 0; k < 7; ++k) fprintf(stderr, " %lld", h[128 + it * 8 + k] - h[it * 16]);
            fprintf(stderr, "\n");
        }
        for (int it = 0; it < 8; ++it) {
            fprintf(stderr, "tc2 timing iter %d:", it);
            for (int k = 0; k < 16; ++k) fprintf(stderr, " %lld", h[it * 16 + k] - h[it * 16]);
            fprintf(stderr, "  | since prev tile start %lld\n", it ? h[it * 16] - h[(it - 1) * 16] : 0);
        }
    }
    return NCA_OK;
}
