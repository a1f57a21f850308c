// ConditionedNCA forward step with the update MLP on the 5th-gen tensor cores (NCA_PREC_BF16).
// Reference semantics: EncoderConditioning/nca.py:152-195 (same step as enc_f32.cu):
//   pre = alive(x); xin = x + goal * pre; p = learned depthwise 3x3 (zero pad); out = Wc relu(Wb relu(Wa p + ba) + bb);
//   x1 = x + fire * out.   The post-update life mask and the clamp stay in enc_life_kernel (enc_f32.cu): they need x1 of the
//   3x3 neighbourhood.
//
// Same skeleton as dynca_tc2.cu: 8x16 tiles (row r = py*16 + px = TMEM lane r), 8 compute warps + 1 MMA / TMA warp, TMA staging
// (zero fill outside the image IS the reference's zero padding, so there are no border patches at all), mbarrier hand-offs.
//   TMA   : x tile + 1-cell ring [C][10][24], living channel + 2-cell ring [12][24], goal tile + ring [C][10][24]
//   pre   : 3x3 max of the living channel > thr at the 10x18 ring positions (thr >= 0: zero fill == -inf padding)
//   xin   : x + goal * pre, written over the goal stage
//   p     : fp32, 4 vertically adjacent cells x 1 channel per thread, 27 learned taps -> A1 [128 x K1] bf16 with
//           k' = 4c + filter (slot 3 of every channel is zero), then the bias chunk [1, 1, 0 ..] (ba as bf16 hi + lo)
//   MMA   : D1 = A1.Wa'^T -> relu -> A2 ; D2 = A2.Wb^T -> + bb, relu -> A3 ; D3 = A3.Wc^T -> x1 = x + fire * D3
#include <mutex>
#include "dynca_tc2.cuh"

#define ET2_NTHREADS 288
#define ET2_HDR 2048u
#define ET2_LR 12          // living stage rows

struct EncTcGeom {
    int B, C, H, W, liv;
    float thr;
    int npairs;            // ceil(C / 2): perception chunks of A1
    int K1;                // (npairs + 1) * 8 rounded up to 16
};

struct EncTcArgs {
    EncTcGeom g;
    float* x1;                           // x + fire * out  [B,C,H,W]
    int slot_in;
    const __nv_bfloat16* Wa; const __nv_bfloat16* Wb; const __nv_bfloat16* Wc;
    const float* bb; const float* wp;    // bb [64], wp [3C][9] fp32
    FireMask fm;
    T2Tiles tl;
};

struct EncTcSmem {
    uint32_t wa, wb, wc, wp, x, l, gl, pre, a1, a2, total;
};
__host__ __device__ static inline EncTcSmem etc_smem(const EncTcGeom& g) {
    EncTcSmem s;
    uint32_t o = ET2_HDR;
    s.wa = o; o += (uint32_t)(g.K1 / 8) * 1024u;
    s.wb = o; o += 8u * 1024u;
    s.wc = o; o += 8u * 512u;
    s.wp = o; o += 64u * 9u * 4u;
    o = (o + 127u) & ~127u;
    s.x = o; o += (uint32_t)g.C * T2_XR * T2_XS * 4u;
    o = (o + 127u) & ~127u;
    s.l = o; o += ET2_LR * T2_XS * 4u;
    o = (o + 127u) & ~127u;
    s.gl = o; o += (uint32_t)g.C * T2_XR * T2_XS * 4u;
    s.pre = o; o += 192u * 4u;
    o = (o + 127u) & ~127u;
    s.a1 = o; o += (uint32_t)(g.K1 / 8) * 2048u;      // A1 (the hidden layers h1 / h2 live in tensor memory)
    s.a2 = o;
    s.total = o;
    return s;
}

// CT: compile-time channel count (20 = the reference's 3 rgb + 1 alpha + 16 hidden; 0 = read from the arguments).  The kernel works
// on a local copy of the geometry with the template value written in, so that loop bounds and shared-memory offsets fold.
template <int CT>
__device__ __forceinline__ void etc_specialize(EncTcGeom& g) {
    if (CT > 0) {
        g.C = CT;
        g.npairs = (CT + 1) / 2;
        g.K1 = ((g.npairs + 1) * 8 + 15) / 16 * 16;
    }
}

template <int CT>
__global__ void __launch_bounds__(ET2_NTHREADS) enc_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                   const __grid_constant__ CUtensorMap tm_l,
                                                                   const __grid_constant__ CUtensorMap tm_g, const EncTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    EncTcGeom g = a.g;
    etc_specialize<CT>(g);
    const EncTcSmem L = etc_smem(g);
    uint64_t* barM = reinterpret_cast<uint64_t*>(smem);
    uint64_t* barT = reinterpret_cast<uint64_t*>(smem + 8);
    uint64_t* barA = reinterpret_cast<uint64_t*>(smem + 16);
    uint64_t* barB = reinterpret_cast<uint64_t*>(smem + 24);
    uint64_t* barC = reinterpret_cast<uint64_t*>(smem + 32);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 40);
    float* sBb = reinterpret_cast<float*>(smem + 64);            // 64 floats
    float* sFire2 = reinterpret_cast<float*>(smem + 512);        // 2 x 128 floats
    uint8_t* sWa = smem + L.wa;
    uint8_t* sWb = smem + L.wb;
    uint8_t* sWc = smem + L.wc;
    float* sWp = reinterpret_cast<float*>(smem + L.wp);
    float* sX = reinterpret_cast<float*>(smem + L.x);
    float* sL = reinterpret_cast<float*>(smem + L.l);
    float* sG = reinterpret_cast<float*>(smem + L.gl);
    float* sPre = reinterpret_cast<float*>(smem + L.pre);
    uint8_t* sA1 = smem + L.a1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = g.C, H = g.H, W = g.W;
    const size_t plane = (size_t)H * W;
    const int n_tiles = a.tl.n_tiles;
    const uint32_t stage_bytes = (uint32_t)(2 * C * T2_XR + ET2_LR) * T2_XS * 4u;

    griddep_launch();

    for (uint32_t i = tid; i < (uint32_t)(g.K1 / 8) * 1024u / 16; i += ET2_NTHREADS)
        reinterpret_cast<uint4*>(sWa)[i] = __ldg(reinterpret_cast<const uint4*>(a.Wa) + i);
    for (uint32_t i = tid; i < 8192u / 16; i += ET2_NTHREADS)
        reinterpret_cast<uint4*>(sWb)[i] = __ldg(reinterpret_cast<const uint4*>(a.Wb) + i);
    for (uint32_t i = tid; i < 4096u / 16; i += ET2_NTHREADS)
        reinterpret_cast<uint4*>(sWc)[i] = __ldg(reinterpret_cast<const uint4*>(a.Wc) + i);
    for (int i = tid; i < 3 * C * 9; i += ET2_NTHREADS) sWp[i] = a.wp[i];
    if (tid < 64) sBb[tid] = a.bb[tid];
    if (tid == 0) {
        mbar_init(barM, 1);
        mbar_init(barT, 1);
        mbar_init(barA, 256);
        mbar_init(barB, 256);
        mbar_init(barC, 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc(tmem_slot, 128u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();          // the previous step's state is complete and visible from here on
    // D1 0..63, D2 64..127, D3 0..31.  The bf16 hidden layers are the A operands of the next MMA straight from tensor memory,
    // written IN PLACE: a thread packs the 32 columns of its lane it has just read into the first 16 of them, so K step ks
    // (16 hidden units) of h1 sits at column 32*(ks/2) + 8*(ks%2), of h2 at 64 + the same.
    const uint32_t TM_D1 = 0u, TM_D2 = 64u, TM_D3 = 0u;

    if (warp == 8) {
        // =========================== MMA / TMA warp ===========================
        const uint32_t id64 = umma_idesc_bf16(128, 64), id32 = umma_idesc_bf16(128, 32);
        const uint64_t dA1 = umma_desc(smem_u32(sA1), 2048u, 128u), dWa = umma_desc(smem_u32(sWa), 1024u, 128u);
        const uint64_t dWb = umma_desc(smem_u32(sWb), 1024u, 128u), dWc = umma_desc(smem_u32(sWc), 512u, 128u);
        const int k1steps = g.K1 / 16;
        const CUtensorMap* const ptm_x = &tm_x;
        const CUtensorMap* const ptm_l = &tm_l;
        const CUtensorMap* const ptm_g = &tm_g;
        uint32_t phA = 0, phB = 0, phC = 0;
        const bool leader = elect_one();
#define ET2_ISSUE_TMA(tile_)                                                                                             \
    do {                                                                                                                 \
        int tb_, ty_, tx_;                                                                                               \
        t2_tile_decode(a.tl, (tile_), tb_, ty_, tx_);                                                                    \
        mbar_expect_tx(barT, stage_bytes);                                                                               \
        tma_load_5d(sX, ptm_x, barT, tx_ - 4, ty_ - 1, 0, tb_, a.slot_in);                                               \
        tma_load_5d(sL, ptm_l, barT, tx_ - 4, ty_ - 2, g.liv < 0 ? 0 : g.liv, tb_, a.slot_in);                           \
        tma_load_5d(sG, ptm_g, barT, tx_ - 4, ty_ - 1, 0, tb_, 0);                                                       \
    } while (0)
        if (leader && (int)blockIdx.x < n_tiles) ET2_ISSUE_TMA(blockIdx.x);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(barA, phA);
            phA ^= 1u;
            tc_fence_after();
            if (leader) {
#pragma unroll 6
                for (int ks = 0; ks < k1steps; ++ks)
                    umma_ss(tmem_base + TM_D1, dA1 + (uint64_t)(ks * (4096 >> 4)), dWa + (uint64_t)(ks * (2048 >> 4)), id64, ks > 0);
                umma_commit(barM);
                if (tile + (int)gridDim.x < n_tiles) ET2_ISSUE_TMA(tile + gridDim.x);     // the stage is consumed
            }
            mbar_wait(barB, phB);
            phB ^= 1u;
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_ts(tmem_base + TM_D2, tmem_base + TM_D1 + 32u * (uint32_t)(ks >> 1) + 8u * (uint32_t)(ks & 1),
                            dWb + (uint64_t)(ks * (2048 >> 4)), id64, ks > 0);
                umma_commit(barM);
            }
            mbar_wait(barC, phC);
            phC ^= 1u;
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_ts(tmem_base + TM_D3, tmem_base + TM_D2 + 32u * (uint32_t)(ks >> 1) + 8u * (uint32_t)(ks & 1),
                            dWc + (uint64_t)(ks * (1024 >> 4)), id32, ks > 0);
                umma_commit(barM);
            }
        }
    } else {
        // =========================== compute warps ===========================
        const int r = tid & 127, half = tid >> 7;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t row_off = (uint32_t)r * 16u;
        const int py = r >> 4, px = r & 15;
        uint32_t phM = 0, phT = 0;
        if (!a.fm.supplied && warp == 7 && (int)blockIdx.x < n_tiles) {
            int tb_, ty_, tx_;
            t2_tile_decode(a.tl, blockIdx.x, tb_, ty_, tx_);
            t2_fire_tile(a.fm, tb_, ty_, tx_, H, W, lane, sFire2, 1);
        }
        bar_sync_n(1, 256);
        int iter = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
            int b, y0, x0;
            t2_tile_decode(a.tl, tile, b, y0, x0);
            const int gy = y0 + py, gx = x0 + px;
            const bool inimg = gy < H && gx < W;
            const float* sFire = sFire2 + (iter & 1) * 128;
            mbar_wait(barT, phT);
            phT ^= 1u;
            // ---- pre-update alive mask at the 10 x 18 ring positions (nca.py:152-163) ----
            if (tid < T2_XR * 18) {
                const int rr = tid / 18, q = tid % 18;
                float mx = 0.0f;                                  // zero fill: harmless for thr >= 0
                const float* lp = sL + rr * T2_XS + T2_XO + q - 1;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) mx = fmaxf(mx, lp[dy * T2_XS + dx]);
                const float pre = (g.liv < 0 || mx > g.thr) ? 1.0f : 0.0f;
                sPre[tid] = pre;
                // ---- xin = x + goal * pre over the goal stage (nca.py:177): this thread's ring position, every channel ----
                const int o0 = rr * T2_XS + T2_XO + q;
#pragma unroll 4
                for (int c = 0; c < C; ++c) {
                    const int o = o0 + c * T2_XR * T2_XS;
                    sG[o] = fmaf(sG[o], pre, sX[o]);
                }
            }
            // residual state of this thread's cell: channels 16*half .. 16*half+15
            float xres[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = 16 * half + i;
                xres[i] = c < C ? sX[(c * T2_XR + py + 1) * T2_XS + T2_XO + px + 1] : 0.0f;
            }
            bar_sync_n(1, 256);
            // ---- learned depthwise 3x3 -> A1: item = (channel pair, 4-row block); lane = (column, channel of the pair) ----
            {
                const int hc = lane & 1, pxx = lane >> 1;      // 8-byte stores of a half warp = 128 contiguous bytes
                for (int item = warp; item < 2 * g.npairs; item += 8) {
                    const int cp = item >> 1, vb = item & 1, c = 2 * cp + hc;
                    float f0[4] = {0.f, 0.f, 0.f, 0.f}, f1[4] = {0.f, 0.f, 0.f, 0.f}, f2[4] = {0.f, 0.f, 0.f, 0.f};
                    if (c < C) {
                        const float* ch = sG + c * T2_XR * T2_XS + (4 * vb) * T2_XS + T2_XO + pxx;
                        const float* w = sWp + c * 27;
                        float v[6][3];
#pragma unroll
                        for (int k = 0; k < 6; ++k)
#pragma unroll
                            for (int j = 0; j < 3; ++j) v[k][j] = ch[k * T2_XS + j];
#pragma unroll
                        for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                            for (int bb = 0; bb < 3; ++bb) {
                                const float w0 = w[aa * 3 + bb], w1 = w[9 + aa * 3 + bb], w2 = w[18 + aa * 3 + bb];
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    f0[k] = fmaf(w0, v[k + aa][bb], f0[k]);
                                    f1[k] = fmaf(w1, v[k + aa][bb], f1[k]);
                                    f2[k] = fmaf(w2, v[k + aa][bb], f2[k]);
                                }
                            }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int rr = (4 * vb + k) * 16 + pxx;
                        uint2 o;
                        o.x = pack_bf16(f0[k], f1[k]); o.y = pack_bf16(f2[k], 0.0f);
                        *reinterpret_cast<uint2*>(sA1 + (uint32_t)cp * 2048u + (uint32_t)rr * 16u + (uint32_t)hc * 8u) = o;
                    }
                }
            }
            if (half == 0) {     // bias chunk [1, 1, 0 ..] and the zero tail chunks
                *reinterpret_cast<uint4*>(sA1 + (uint32_t)g.npairs * 2048u + row_off) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
            } else {
                for (int ch = g.npairs + 1; ch < g.K1 / 8; ++ch)
                    *reinterpret_cast<uint4*>(sA1 + (uint32_t)ch * 2048u + row_off) = make_uint4(0, 0, 0, 0);
            }
            // fire decisions of the NEXT tile (double buffered)
            if (!a.fm.supplied && warp == 7 && tile + (int)gridDim.x < n_tiles) {
                int tb_, ty_, tx_;
                t2_tile_decode(a.tl, tile + gridDim.x, tb_, ty_, tx_);
                t2_fire_tile(a.fm, tb_, ty_, tx_, H, W, lane, sFire2 + ((iter + 1) & 1) * 128, 1);
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(barA);
            mbar_wait(barM, phM);
            phM ^= 1u;
            tc_fence_after();
            // ---- E1: h1 = relu(D1) -> bf16, in place in tensor memory (ba rides in the bias chunk) ----
            {
                uint32_t v[32], o[16];
                tmem_ld32(tmem_lane + TM_D1 + 32u * (uint32_t)half, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = pack_bf16_relu(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                tmem_st16(tmem_lane + TM_D1 + 32u * (uint32_t)half, o);
                tmem_st_wait();
            }
            tc_fence_before();
            mbar_arrive(barB);
            mbar_wait(barM, phM);
            phM ^= 1u;
            tc_fence_after();
            // ---- E2: h2 = relu(D2 + bb) -> bf16, in place in tensor memory ----
            {
                uint32_t v[32], o[16];
                tmem_ld32(tmem_lane + TM_D2 + 32u * (uint32_t)half, v);
                tmem_ld_wait();
                const float* bbp = sBb + 32 * half;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    o[i] = pack_bf16_relu(__uint_as_float(v[2 * i]) + bbp[2 * i], __uint_as_float(v[2 * i + 1]) + bbp[2 * i + 1]);
                tmem_st16(tmem_lane + TM_D2 + 32u * (uint32_t)half, o);
                tmem_st_wait();
            }
            tc_fence_before();
            mbar_arrive(barC);
            mbar_wait(barM, phM);
            phM ^= 1u;
            tc_fence_after();
            // ---- E3: x1 = x + fire * D3 (nca.py:186) ----
            {
                uint32_t v[16];
                tmem_ld16(tmem_lane + TM_D3 + 16u * (uint32_t)half, v);
                tmem_ld_wait();
                const float fire = a.fm.supplied ? (inimg ? a.fm.supplied[((size_t)b * H + gy) * W + gx] : 0.0f) : sFire[r];
                const int nch = min(16, C - 16 * half);                // warp-uniform
                float* xo = a.x1 + ((size_t)b * C + 16 * half) * plane + (size_t)gy * W + gx;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (inimg && i < nch) *xo = fmaf(fire, __uint_as_float(v[i]), xres[i]);
                    xo += plane;
                }
            }
            tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 128u);
}

// operand images: Wa' [N = 64][K = K1] (k' = 4c + f; bias columns 4C' .. as bf16 hi + lo), Wb [64][64], Wc [N = 32][K = 64]
__global__ void enc_tc_prep_kernel(EncTcGeom g, const float* __restrict__ wa, const float* __restrict__ ba,
                                   const float* __restrict__ wb, const float* __restrict__ wc, __nv_bfloat16* __restrict__ Wa,
                                   __nv_bfloat16* __restrict__ Wb, __nv_bfloat16* __restrict__ Wc) {
    const int n1 = g.K1 * 64, n2 = 64 * 64, n3 = 64 * 32;
    const int K = 3 * g.C;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + n3; i += gridDim.x * blockDim.x) {
        if (i < n1) {
            const int kp = i / 64, j = i % 64;           // (k', hidden unit)
            const int kc = kp >> 3, s = kp & 7;
            float v = 0.0f;
            if (kc < g.npairs) {
                const int c = 2 * kc + (s >> 2), f = s & 3;
                if (c < g.C && f < 3) v = wa[j * K + 3 * c + f];
            } else if (kc == g.npairs) {
                const float hi = __bfloat162float(__float2bfloat16_rn(ba[j]));
                if (s == 0) v = hi; else if (s == 1) v = ba[j] - hi;
            }
            Wa[(size_t)kc * 512 + (size_t)j * 8 + s] = __float2bfloat16_rn(v);
        } else if (i < n1 + n2) {
            const int e = i - n1, k = e / 64, j = e % 64;     // h2[j] = sum_k wb[j][k] h1[k]
            Wb[(size_t)(k >> 3) * 512 + (size_t)j * 8 + (k & 7)] = __float2bfloat16_rn(wb[j * 64 + k]);
        } else {
            const int e = i - n1 - n2, k = e / 32, c = e % 32;
            Wc[(size_t)(k >> 3) * 256 + (size_t)c * 8 + (k & 7)] = __float2bfloat16_rn(c < g.C ? wc[c * 64 + k] : 0.0f);
        }
    }
}

// ---- host side --------------------------------------------------------------------------------------------
static int etc_make_geom(const NcaEncDesc* d, EncTcGeom* g) {
    g->B = d->B; g->C = d->C; g->H = d->H; g->W = d->W; g->liv = d->living_dim < 0 ? -1 : d->living_dim;
    g->thr = d->alive_thr;
    g->npairs = (d->C + 1) / 2;
    g->K1 = ((g->npairs + 1) * 8 + 15) / 16 * 16;
    return NCA_OK;
}

bool enc_tc_supported(const NcaEncDesc* d) {
    if (d->precision != NCA_PREC_BF16) return false;
    if (d->hid != 64 || d->C > 22 || d->C < 1 || (d->W & 3) != 0 || d->alive_thr < 0.0f) return false;
    if ((long long)d->H * d->W >= (1ll << 30)) return false;
    EncTcGeom g;
    etc_make_geom(d, &g);
    return etc_smem(g).total <= 227u * 1024u;
}

size_t enc_tc_weight_bytes(const NcaEncDesc* d) {
    EncTcGeom g;
    etc_make_geom(d, &g);
    return nca_align_up((size_t)g.K1 * 64 * 2 + 64 * 64 * 2 + 64 * 32 * 2, 256);
}

// one forward step: x1 = x + fire * update(x, goal).  ws: enc_tc_weight_bytes; maps are built per call by enc_tc_make_maps
int enc_tc_prep_weights(const NcaEncDesc* d, const NcaEncWeights* w, void* ws, cudaStream_t s) {
    EncTcGeom g;
    etc_make_geom(d, &g);
    __nv_bfloat16* Wa = (__nv_bfloat16*)ws;
    __nv_bfloat16* Wb = Wa + (size_t)g.K1 * 64;
    __nv_bfloat16* Wc = Wb + 64 * 64;
    enc_tc_prep_kernel<<<16, 256, 0, s>>>(g, w->wa, w->ba, w->wb, w->wc, Wa, Wb, Wc);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int enc_tc_make_maps(const NcaEncDesc* d, const float* states, int slots, const float* goal, EncTcMaps* m) {
    const size_t n = (size_t)d->B * d->C * d->H * d->W;
    int rc = t2_make_map((CUtensorMap*)m->x, states, slots, n, d->B, d->C, d->H, d->W, T2_XR, T2_XS);
    if (rc) return rc;
    rc = t2_make_map((CUtensorMap*)m->l, states, slots, n, d->B, d->C, d->H, d->W, ET2_LR, T2_XS, 1);
    if (rc) return rc;
    return t2_make_map((CUtensorMap*)m->g, goal, 1, n, d->B, d->C, d->H, d->W, T2_XR, T2_XS);
}

int enc_tc_forward_step(const NcaEncDesc* d, const NcaEncWeights* w, const void* ws, const EncTcMaps* m, int slot_in, float* x1,
                        const FireMask& fm, cudaStream_t s, int pdl) {
    EncTcArgs a;
    etc_make_geom(d, &a.g);
    a.x1 = x1; a.slot_in = slot_in;
    a.Wa = (const __nv_bfloat16*)ws;
    a.Wb = a.Wa + (size_t)a.g.K1 * 64;
    a.Wc = a.Wb + 64 * 64;
    a.bb = w->bb; a.wp = w->wp;
    a.fm = fm;
    a.tl = t2_make_tiles(d->B, d->H, d->W);
    const size_t smem = etc_smem(a.g).total;
    // co-resident CTAs: registers, shared memory (once per shared-memory size), TMEM (128 columns each)
    static std::mutex occ_mu;
    static size_t occ_smem_dev[NCA_MAX_DEVICES][2];
    static int occ_val_dev[NCA_MAX_DEVICES][2];
    const int dev = nca_device_ordinal() % NCA_MAX_DEVICES;
    const int spec = d->C == 20 ? 1 : 0;           // the reference's channel count has its own instantiation
    std::unique_lock<std::mutex> occ_lock(occ_mu);
    size_t& occ_smem = occ_smem_dev[dev][spec];
    int& occ_val = occ_val_dev[dev][spec];
    if (occ_val == 0 || occ_smem != smem) {
        int o = 0;
        if (spec) {
            NCA_CUDA_OK(cudaFuncSetAttribute(enc_fwd_tc_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            o = t2_occupancy_by_regs(enc_fwd_tc_kernel<20>, ET2_NTHREADS);
        } else {
        NCA_CUDA_OK(cudaFuncSetAttribute(enc_fwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        o = t2_occupancy_by_regs(enc_fwd_tc_kernel<0>, ET2_NTHREADS);
        }
        const int by_smem = (int)((227 * 1024) / (smem + 1024));
        if (o > by_smem) o = by_smem;
        occ_val = o < 1 ? 1 : o;
        occ_smem = smem;
    }
    int occ = occ_val;
    occ_lock.unlock();
    if (occ > 4) occ = 4;
    if (occ < 1) occ = 1;
    int grid = t2_num_sms() * occ;
    if (grid > a.tl.n_tiles) grid = a.tl.n_tiles;
    if (spec)
        NCA_CUDA_OK(t2_launch(enc_fwd_tc_kernel<20>, grid, ET2_NTHREADS, smem, s, pdl != 0, *(const CUtensorMap*)m->x, *(const CUtensorMap*)m->l,
                              *(const CUtensorMap*)m->g, a));
    else
        NCA_CUDA_OK(t2_launch(enc_fwd_tc_kernel<0>, grid, ET2_NTHREADS, smem, s, pdl != 0, *(const CUtensorMap*)m->x, *(const CUtensorMap*)m->l,
                              *(const CUtensorMap*)m->g, a));
    NCA_LAUNCH_OK();
    return NCA_OK;
}

// =====================================================================================================================
// ConditionedNCA BPTT step on tcgen05.  One CTA per SM: 16 compute warps (512 threads, one 8x16 tile at a time) + 1 MMA /
// TMA warp; thread (r = tid & 127, q = tid >> 7) owns cell r (TMEM lane r).
//   recompute : p -> A1, D1 = A1.Wa'^T, h1 -> H1, D2 = H1.Wb^T, h2 -> H2, D3 = H2.Wc^T, x1 = x + fire * D3
//   E3        : g1 = g * [|x1 * life| <= clamp] * life (clamp, then life mask);  g_o = fire * g1 -> Go
//   dgrad     : D4 = Go.Wc -> g_a2 = D4 * [h2 > 0] ; D5 = Ga2.Wb -> g_a1 = D5 * [h1 > 0] ; D6 = Ga1.Wa' (g_p)
//   wgrad     : D7 += H2^T.Go (gWc^T) ; D8 += Ga2^T.H1 (gWb) ; D10 += Ga2^T.1 (gbb) ; D9 += Ga1^T.A1 (gWa', gba)  -- in TMEM
//               over all tiles of the CTA (cells are K through MN-major views), flushed once with red.add
//   conv^T    : g_p -> zero-padded fp32 planes -> 27-tap transposed learned depthwise conv (5 vertical outputs per thread)
//               -> dL/dx_t (red.add, + g1 pass-through) and dL/dgoal (red.add where pre-alive)
//   gwp       : per-thread register accumulators (warps 0..9 = channel pairs), reduced once at the end
// TMEM columns: [0,96) D1 -> D4 -> D6 | [96,160) D2 -> D5 | [160,192) D3 | [192,224) D7 | [224,288) D8 | [288,384) D9 | [384,400) D10
// =====================================================================================================================
#define EB_NCOMP 512
#define EB_NTHREADS 544
#define EB_PR 12
#define EB_PS 20
#define EB_PP (EB_PR * EB_PS)

struct EncTcBwdArgs {
    EncTcGeom g;
    float clampv;
    int slot_in;
    const uint8_t* life;                 // [B,H,W] life mask of this step
    float* g_out; float* g_goal;         // red.add targets
    const __nv_bfloat16* Wa; const __nv_bfloat16* Wb; const __nv_bfloat16* Wc;
    const float* bb; const float* wp;
    float* gwp; float* gwa; float* gba; float* gwb; float* gbb; float* gwc;     // reference layouts (red.add)
    FireMask fm;
    T2Tiles tl;
};

struct EncTcBwdSmem {
    uint32_t wa, wb, wc, ones, wp, x, l, gl, gn, pre, xin, a1, h1, h2, go, ga2, ga1, pl, base, total;
};
__host__ __device__ static inline EncTcBwdSmem etb_smem(const EncTcGeom& g) {
    EncTcBwdSmem s;
    const uint32_t C = (uint32_t)g.C;
    uint32_t o = ET2_HDR;
    s.wa = o; o += (uint32_t)(g.K1 / 8) * 1024u;
    s.wb = o; o += 8u * 1024u;
    s.wc = o; o += 8u * 512u;
    s.ones = o; o += 2u * 2048u;                      // [cells][16] bf16 ones, MN-major B operand of the gbb MMA
    s.wp = o; o += 64u * 9u * 4u;
    o = (o + 127u) & ~127u;
    s.x = o; o += C * T2_XR * T2_XS * 4u;
    o = (o + 127u) & ~127u;
    s.l = o; o += ET2_LR * T2_XS * 4u;
    o = (o + 127u) & ~127u;
    s.gl = o; o += C * T2_XR * T2_XS * 4u;
    o = (o + 127u) & ~127u;
    s.gn = o; o += C * T2_TH * T2_TW * 4u;
    s.pre = o; o += 192u * 4u;
    s.xin = o; o += C * T2_XR * 18u * 4u;             // [C][10][18]
    o = (o + 127u) & ~127u;
    // ---- operand region (overlaid by the planes once every MMA of the tile is complete) ----
    const uint32_t base = o;
    s.a1 = o; o += (uint32_t)(g.K1 / 8) * 2048u;
    s.go = o; o += 4u * 2048u;
    s.h1 = o; o += 8u * 2048u;
    s.h2 = o; o += 8u * 2048u;                        // the M = 128 views of H2 / Ga2 / Ga1 read 8 chunks past their end:
    s.ga2 = o; o += 8u * 2048u;                       //   they are followed by allocated memory (the last one by the slack)
    s.ga1 = o; o += 8u * 2048u;
    o += 8u * 2048u;                                  // slack for the view of Ga1
    uint32_t p = base;
    s.pl = p; p += 3u * C * EB_PP * 4u;
    s.base = p; p += C * T2_TH * T2_TW * 4u;
    s.total = o > p ? o : p;
    return s;
}

// transposed learned depthwise conv at ring position (oy0 + k, ox), k < NR: sum_f sum_{a,b} w[f][a][b] P_f[cell(oy - a, ox - b)]
// planes P_f: zero padded [12][20], cell (py,px) at [py+2][px+2]; w = 27 weights of the channel
template <int NR>
__device__ __forceinline__ void eb_conv_t(const float* __restrict__ P0, const float* __restrict__ w, int oy0, int ox, float out[NR]) {
#pragma unroll
    for (int k = 0; k < NR; ++k) out[k] = 0.0f;
#pragma unroll
    for (int f = 0; f < 3; ++f) {
        const float* P = P0 + f * EB_PP;
        float v[NR + 2][3];
#pragma unroll
        for (int k = 0; k < NR + 2; ++k)
#pragma unroll
            for (int j = 0; j < 3; ++j) v[k][j] = P[(oy0 + k) * EB_PS + ox + j];      // cell row oy0 + k - 2, column ox + j - 2
#pragma unroll
        for (int aa = 0; aa < 3; ++aa)
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) {
                const float ww = w[f * 9 + aa * 3 + bb];
#pragma unroll
                for (int k = 0; k < NR; ++k) out[k] = fmaf(ww, v[k + 2 - aa][2 - bb], out[k]);   // cell (oy - a, ox - b)
            }
    }
}

template <int CT>
__global__ void __launch_bounds__(EB_NTHREADS, 1) enc_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                     const __grid_constant__ CUtensorMap tm_l,
                                                                     const __grid_constant__ CUtensorMap tm_g,
                                                                     const __grid_constant__ CUtensorMap tm_gn, const EncTcBwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    EncTcGeom g = a.g;
    etc_specialize<CT>(g);
    const EncTcBwdSmem L = etb_smem(g);
    uint64_t* barM = reinterpret_cast<uint64_t*>(smem);
    uint64_t* barT = reinterpret_cast<uint64_t*>(smem + 8);
    uint64_t* barH = reinterpret_cast<uint64_t*>(smem + 16);      // 6 hand-off barriers A..F
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
    float* sBb = reinterpret_cast<float*>(smem + 128);            // 64 floats
    float* sFire2 = reinterpret_cast<float*>(smem + 512);         // 2 x 128 floats
    uint8_t* sWa = smem + L.wa;
    uint8_t* sWb = smem + L.wb;
    uint8_t* sWc = smem + L.wc;
    uint8_t* sOnes = smem + L.ones;
    float* sWp = reinterpret_cast<float*>(smem + L.wp);
    float* sX = reinterpret_cast<float*>(smem + L.x);
    float* sL = reinterpret_cast<float*>(smem + L.l);
    float* sG = reinterpret_cast<float*>(smem + L.gl);
    float* sGn = reinterpret_cast<float*>(smem + L.gn);
    float* sPre = reinterpret_cast<float*>(smem + L.pre);
    float* sXin = reinterpret_cast<float*>(smem + L.xin);
    uint8_t* sA1 = smem + L.a1;
    uint8_t* sGo = smem + L.go;
    uint8_t* sH1 = smem + L.h1;
    uint8_t* sH2 = smem + L.h2;
    uint8_t* sGa2 = smem + L.ga2;
    uint8_t* sGa1 = smem + L.ga1;
    float* sPl = reinterpret_cast<float*>(smem + L.pl);           // [C][3][12][20]
    float* sBase = reinterpret_cast<float*>(smem + L.base);       // [C][8][16]: g1 (pass-through)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = g.C, H = g.H, W = g.W;
    const size_t plane = (size_t)H * W;
    const int n_tiles = a.tl.n_tiles;
    const uint32_t stage_bytes = (uint32_t)(2 * C * T2_XR + ET2_LR) * T2_XS * 4u + (uint32_t)C * T2_TH * T2_TW * 4u;

    for (uint32_t i = tid; i < (uint32_t)(g.K1 / 8) * 1024u / 16; i += EB_NTHREADS)
        reinterpret_cast<uint4*>(sWa)[i] = __ldg(reinterpret_cast<const uint4*>(a.Wa) + i);
    for (uint32_t i = tid; i < 8192u / 16; i += EB_NTHREADS)
        reinterpret_cast<uint4*>(sWb)[i] = __ldg(reinterpret_cast<const uint4*>(a.Wb) + i);
    for (uint32_t i = tid; i < 4096u / 16; i += EB_NTHREADS)
        reinterpret_cast<uint4*>(sWc)[i] = __ldg(reinterpret_cast<const uint4*>(a.Wc) + i);
    for (uint32_t i = tid; i < 4096u / 16; i += EB_NTHREADS)
        reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    for (int i = tid; i < 3 * C * 9; i += EB_NTHREADS) sWp[i] = a.wp[i];
    if (tid < 64) sBb[tid] = a.bb[tid];
    // the operand region must be finite before the first MMAs read rows nobody writes
    for (uint32_t i = L.a1 / 16 + tid; i < L.total / 16; i += EB_NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(barM, 1);
        mbar_init(barT, 1);
        for (int i = 0; i < 6; ++i) mbar_init(barH + i, EB_NCOMP);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) tmem_alloc(tmem_slot, 512u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t TM_A = 0u, TM_B = 96u, TM_C = 160u, TM_D7 = 192u, TM_D8 = 224u, TM_D9 = 288u, TM_D10 = 384u;

    if (warp == 16) {
        // =========================== MMA / TMA warp ===========================
        const uint32_t id64 = umma_idesc_bf16(128, 64), id32 = umma_idesc_bf16(128, 32);
        const uint32_t id64_bmn = id64 | (1u << 16), idK1_bmn = umma_idesc_bf16(128, g.K1) | (1u << 16);
        const uint32_t id32_mn = umma_idesc_bf16_mn(128, 32), id64_mn = umma_idesc_bf16_mn(128, 64);
        const uint32_t id16_mn = umma_idesc_bf16_mn(128, 16), idK1_mn = umma_idesc_bf16_mn(128, g.K1);
        const uint64_t dA1 = umma_desc(smem_u32(sA1), 2048u, 128u), dWa = umma_desc(smem_u32(sWa), 1024u, 128u);
        const uint64_t dH1 = umma_desc(smem_u32(sH1), 2048u, 128u), dWb = umma_desc(smem_u32(sWb), 1024u, 128u);
        const uint64_t dH2 = umma_desc(smem_u32(sH2), 2048u, 128u), dWc = umma_desc(smem_u32(sWc), 512u, 128u);
        const uint64_t dGo = umma_desc(smem_u32(sGo), 2048u, 128u), dGa2 = umma_desc(smem_u32(sGa2), 2048u, 128u);
        const uint64_t dGa1 = umma_desc(smem_u32(sGa1), 2048u, 128u);
        // weights viewed MN-major (the K index of the forward GEMM becomes N): LBO = 128, SBO = chunk stride
        const uint64_t dWcT = umma_desc(smem_u32(sWc), 128u, 512u), dWbT = umma_desc(smem_u32(sWb), 128u, 1024u);
        const uint64_t dWaT = umma_desc(smem_u32(sWa), 128u, 1024u);
        // activations viewed MN-major (cells become K)
        const uint64_t dH2t = umma_desc(smem_u32(sH2), 128u, 2048u), dGot = umma_desc(smem_u32(sGo), 128u, 2048u);
        const uint64_t dGa2t = umma_desc(smem_u32(sGa2), 128u, 2048u), dH1t = umma_desc(smem_u32(sH1), 128u, 2048u);
        const uint64_t dGa1t = umma_desc(smem_u32(sGa1), 128u, 2048u), dA1t = umma_desc(smem_u32(sA1), 128u, 2048u);
        const uint64_t dOnes = umma_desc(smem_u32(sOnes), 128u, 2048u);
        const int k1steps = g.K1 / 16;
        const CUtensorMap* const ptm_x = &tm_x;
        const CUtensorMap* const ptm_l = &tm_l;
        const CUtensorMap* const ptm_g = &tm_g;
        const CUtensorMap* const ptm_gn = &tm_gn;
        uint32_t ph[6] = {0, 0, 0, 0, 0, 0};
        const bool leader = elect_one();
        bool first = true;
#define EB_ISSUE_TMA(tile_)                                                                                              \
    do {                                                                                                                 \
        int tb_, ty_, tx_;                                                                                               \
        t2_tile_decode(a.tl, (tile_), tb_, ty_, tx_);                                                                    \
        mbar_expect_tx(barT, stage_bytes);                                                                               \
        tma_load_5d(sX, ptm_x, barT, tx_ - 4, ty_ - 1, 0, tb_, a.slot_in);                                               \
        tma_load_5d(sL, ptm_l, barT, tx_ - 4, ty_ - 2, g.liv < 0 ? 0 : g.liv, tb_, a.slot_in);                           \
        tma_load_5d(sG, ptm_g, barT, tx_ - 4, ty_ - 1, 0, tb_, 0);                                                       \
        tma_load_5d(sGn, ptm_gn, barT, tx_, ty_, 0, tb_, 0);                                                             \
    } while (0)
#define EB_WAIT(i_) do { mbar_wait(barH + (i_), ph[i_]); ph[i_] ^= 1u; tc_fence_after(); } while (0)
        if (leader && (int)blockIdx.x < n_tiles) EB_ISSUE_TMA(blockIdx.x);
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            EB_WAIT(0);                                        // A1 written, stage consumed
            if (leader) {
#pragma unroll 6
                for (int ks = 0; ks < k1steps; ++ks)
                    umma_ss(tmem_base + TM_A, dA1 + (uint64_t)(ks * (4096 >> 4)), dWa + (uint64_t)(ks * (2048 >> 4)), id64, ks > 0);
                umma_commit(barM);
                if (tile + (int)gridDim.x < n_tiles) EB_ISSUE_TMA(tile + gridDim.x);
            }
            EB_WAIT(1);                                        // H1
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_ss(tmem_base + TM_B, dH1 + (uint64_t)(ks * (4096 >> 4)), dWb + (uint64_t)(ks * (2048 >> 4)), id64, ks > 0);
                umma_commit(barM);
            }
            EB_WAIT(2);                                        // H2
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_ss(tmem_base + TM_C, dH2 + (uint64_t)(ks * (4096 >> 4)), dWc + (uint64_t)(ks * (1024 >> 4)), id32, ks > 0);
                umma_commit(barM);
            }
            EB_WAIT(3);                                        // Go
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)                 // D4 = Go . Wc  (K = 32 channels)
                    umma_ss(tmem_base + TM_A, dGo + (uint64_t)(ks * (4096 >> 4)), dWcT + (uint64_t)(ks * (256 >> 4)), id64_bmn, ks > 0);
                umma_commit(barM);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {               // D7 += H2^T . Go   [hidden x channel]
                    const uint64_t o = (uint64_t)(ks * (256 >> 4));
                    umma_ss(tmem_base + TM_D7, dH2t + o, dGot + o, id32_mn, !(first && ks == 0));
                }
            }
            EB_WAIT(4);                                        // Ga2
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)                 // D5 = Ga2 . Wb
                    umma_ss(tmem_base + TM_B, dGa2 + (uint64_t)(ks * (4096 >> 4)), dWbT + (uint64_t)(ks * (256 >> 4)), id64_bmn, ks > 0);
                umma_commit(barM);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {               // D8 += Ga2^T . H1 ; D10 += Ga2^T . 1
                    const uint64_t o = (uint64_t)(ks * (256 >> 4));
                    umma_ss(tmem_base + TM_D8, dGa2t + o, dH1t + o, id64_mn, !(first && ks == 0));
                    umma_ss(tmem_base + TM_D10, dGa2t + o, dOnes + o, id16_mn, !(first && ks == 0));
                }
            }
            EB_WAIT(5);                                        // Ga1
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)                 // D6 = Ga1 . Wa'  (g_p, K1 columns)
                    umma_ss(tmem_base + TM_A, dGa1 + (uint64_t)(ks * (4096 >> 4)), dWaT + (uint64_t)(ks * (256 >> 4)), idK1_bmn, ks > 0);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {               // D9 += Ga1^T . A1
                    const uint64_t o = (uint64_t)(ks * (256 >> 4));
                    umma_ss(tmem_base + TM_D9, dGa1t + o, dA1t + o, idK1_mn, !(first && ks == 0));
                }
                umma_commit(barM);                             // everything of this tile is complete
            }
            first = false;
        }
    } else {
        // =========================== compute warps ===========================
        const int r = tid & 127, qtr = tid >> 7;
        const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t row_off = (uint32_t)r * 16u;
        const int py = r >> 4, px = r & 15;
        uint32_t phM = 0, phT = 0;
        float wpacc[27];
#pragma unroll
        for (int i = 0; i < 27; ++i) wpacc[i] = 0.0f;
        // zero ring of the planes: 3C planes x 28 items (rows 0,1,10,11 as float4; columns 0,1,18,19 of rows 2..9)
        int zoff[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = tid + q * EB_NCOMP;
            int o = -1;
            if (i < 3 * C * 28) {
                const int k = i % 28, pl = i / 28;
                if (k < 20) { const int row = k / 5 < 2 ? k / 5 : 8 + k / 5; o = pl * EB_PP + row * EB_PS + 4 * (k % 5); }
                else o = (pl * EB_PP + (2 + (k - 20)) * EB_PS) | (1 << 30);
            }
            zoff[q] = o;
        }
        if (!a.fm.supplied && warp == 15 && (int)blockIdx.x < n_tiles) {
            int tb_, ty_, tx_;
            t2_tile_decode(a.tl, blockIdx.x, tb_, ty_, tx_);
            t2_fire_tile(a.fm, tb_, ty_, tx_, H, W, lane, sFire2, 1);
        }
        bar_sync_n(1, EB_NCOMP);
#define EB_ARRIVE(i_) do { fence_proxy_async(); tc_fence_before(); mbar_arrive(barH + (i_)); } while (0)
#define EB_WAITM() do { mbar_wait(barM, phM); phM ^= 1u; tc_fence_after(); } while (0)
        int iter = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++iter) {
            int b, y0, x0;
            t2_tile_decode(a.tl, tile, b, y0, x0);
            const int gy = y0 + py, gx = x0 + px;
            const bool inimg = gy < H && gx < W;
            const float* sFire = sFire2 + (iter & 1) * 128;
            mbar_wait(barT, phT);
            phT ^= 1u;
            // ---- pre-update alive mask at the ring positions; this thread's g and x (channels 8q .. 8q+7) ----
            if (tid < 2 * T2_XR * 18) {      // two threads per ring position: even / odd channels of xin = x + goal * pre (nca.py:177)
                const int grp = tid >= T2_XR * 18 ? 1 : 0, p = tid - grp * T2_XR * 18;
                const int rr = p / 18, q = p % 18;
                float mx = 0.0f;
                const float* lp = sL + rr * T2_XS + T2_XO + q - 1;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) mx = fmaxf(mx, lp[dy * T2_XS + dx]);
                const float pre = (g.liv < 0 || mx > g.thr) ? 1.0f : 0.0f;
                if (grp == 0) sPre[p] = pre;
                const int o0 = rr * T2_XS + T2_XO + q;
#pragma unroll 2
                for (int c = grp; c < C; c += 2) {
                    const int o = o0 + c * T2_XR * T2_XS;
                    sXin[(c * T2_XR + rr) * 18 + q] = fmaf(sG[o], pre, sX[o]);
                }
            }
            float xres[8], gn[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = 8 * qtr + i;
                xres[i] = c < C ? sX[(c * T2_XR + py + 1) * T2_XS + T2_XO + px + 1] : 0.0f;
                gn[i] = c < C ? sGn[(c * T2_TH + py) * T2_TW + px] : 0.0f;
            }
            bar_sync_n(1, EB_NCOMP);
            // ---- learned depthwise 3x3 -> A1 ----
            {
                const int hc = lane & 1, pxx = lane >> 1;      // 8-byte stores of a half warp = 128 contiguous bytes
                for (int item = warp; item < 2 * g.npairs; item += 16) {
                    const int cp = item >> 1, vb = item & 1, c = 2 * cp + hc;
                    float f0[4] = {0.f, 0.f, 0.f, 0.f}, f1[4] = {0.f, 0.f, 0.f, 0.f}, f2[4] = {0.f, 0.f, 0.f, 0.f};
                    if (c < C) {
                        const float* ch = sXin + (c * T2_XR + 4 * vb) * 18 + pxx;
                        const float* w = sWp + c * 27;
                        float v[6][3];
#pragma unroll
                        for (int k = 0; k < 6; ++k)
#pragma unroll
                            for (int j = 0; j < 3; ++j) v[k][j] = ch[k * 18 + j];
#pragma unroll
                        for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                            for (int bb = 0; bb < 3; ++bb) {
                                const float w0 = w[aa * 3 + bb], w1 = w[9 + aa * 3 + bb], w2 = w[18 + aa * 3 + bb];
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    f0[k] = fmaf(w0, v[k + aa][bb], f0[k]);
                                    f1[k] = fmaf(w1, v[k + aa][bb], f1[k]);
                                    f2[k] = fmaf(w2, v[k + aa][bb], f2[k]);
                                }
                            }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int rr = (4 * vb + k) * 16 + pxx;
                        uint2 o;
                        o.x = pack_bf16(f0[k], f1[k]); o.y = pack_bf16(f2[k], 0.0f);
                        *reinterpret_cast<uint2*>(sA1 + (uint32_t)cp * 2048u + (uint32_t)rr * 16u + (uint32_t)hc * 8u) = o;
                    }
                }
            }
            if (qtr == 0) {      // bias chunk (zero for cells outside the image: A1 rows are summed over cells by D9)
                *reinterpret_cast<uint4*>(sA1 + (uint32_t)g.npairs * 2048u + row_off) = make_uint4(inimg ? 0x3F803F80u : 0u, 0u, 0u, 0u);
            } else if (qtr == 1) {
                for (int ch = g.npairs + 1; ch < g.K1 / 8; ++ch)
                    *reinterpret_cast<uint4*>(sA1 + (uint32_t)ch * 2048u + row_off) = make_uint4(0, 0, 0, 0);
            }
            if (!a.fm.supplied && warp == 15 && tile + (int)gridDim.x < n_tiles) {
                int tb_, ty_, tx_;
                t2_tile_decode(a.tl, tile + gridDim.x, tb_, ty_, tx_);
                t2_fire_tile(a.fm, tb_, ty_, tx_, H, W, lane, sFire2 + ((iter + 1) & 1) * 128, 1);
            }
            EB_ARRIVE(0);
            EB_WAITM();
            // ---- E1: h1 = relu(D1) -> H1; thread -> hidden units 16q .. 16q+15 ----
            {
                uint32_t v[16];
                tmem_ld16(tmem_lane + TM_A + 16u * (uint32_t)qtr, v);
                tmem_ld_wait();
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {
                    uint4 o;
                    o.x = pack_bf16_relu(__uint_as_float(v[qq * 8 + 0]), __uint_as_float(v[qq * 8 + 1]));
                    o.y = pack_bf16_relu(__uint_as_float(v[qq * 8 + 2]), __uint_as_float(v[qq * 8 + 3]));
                    o.z = pack_bf16_relu(__uint_as_float(v[qq * 8 + 4]), __uint_as_float(v[qq * 8 + 5]));
                    o.w = pack_bf16_relu(__uint_as_float(v[qq * 8 + 6]), __uint_as_float(v[qq * 8 + 7]));
                    *reinterpret_cast<uint4*>(sH1 + (uint32_t)(2 * qtr + qq) * 2048u + row_off) = o;
                }
            }
            EB_ARRIVE(1);
            EB_WAITM();
            // ---- E2: h2 = relu(D2 + bb) -> H2 ----
            {
                uint32_t v[16];
                tmem_ld16(tmem_lane + TM_B + 16u * (uint32_t)qtr, v);
                tmem_ld_wait();
                const float* bbp = sBb + 16 * qtr;
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {
                    uint4 o;
                    o.x = pack_bf16_relu(__uint_as_float(v[qq * 8 + 0]) + bbp[qq * 8 + 0], __uint_as_float(v[qq * 8 + 1]) + bbp[qq * 8 + 1]);
                    o.y = pack_bf16_relu(__uint_as_float(v[qq * 8 + 2]) + bbp[qq * 8 + 2], __uint_as_float(v[qq * 8 + 3]) + bbp[qq * 8 + 3]);
                    o.z = pack_bf16_relu(__uint_as_float(v[qq * 8 + 4]) + bbp[qq * 8 + 4], __uint_as_float(v[qq * 8 + 5]) + bbp[qq * 8 + 5]);
                    o.w = pack_bf16_relu(__uint_as_float(v[qq * 8 + 6]) + bbp[qq * 8 + 6], __uint_as_float(v[qq * 8 + 7]) + bbp[qq * 8 + 7]);
                    *reinterpret_cast<uint4*>(sH2 + (uint32_t)(2 * qtr + qq) * 2048u + row_off) = o;
                }
            }
            EB_ARRIVE(2);
            EB_WAITM();
            // ---- E3: x1 = x + fire * D3; g1 = g * [|x1 * life| <= clamp] * life; g_o = fire * g1 -> Go (nca.py:186-194) ----
            float g1[8];
            {
                uint32_t v[8];
                tmem_ld8(tmem_lane + TM_C + 8u * (uint32_t)qtr, v);
                tmem_ld_wait();
                const float fire = a.fm.supplied ? (inimg ? a.fm.supplied[((size_t)b * H + gy) * W + gx] : 0.0f) : sFire[r];
                const float life = inimg ? (float)a.life[((size_t)b * H + gy) * W + gx] : 0.0f;
                float go[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float x1 = fmaf(fire, __uint_as_float(v[i]), xres[i]);
                    const float vv = x1 * life;
                    g1[i] = (8 * qtr + i < C && vv >= -a.clampv && vv <= a.clampv) ? gn[i] * life : 0.0f;
                    go[i] = fire * g1[i];
                }
                uint4 o;
                o.x = pack_bf16(go[0], go[1]); o.y = pack_bf16(go[2], go[3]); o.z = pack_bf16(go[4], go[5]); o.w = pack_bf16(go[6], go[7]);
                *reinterpret_cast<uint4*>(sGo + (uint32_t)qtr * 2048u + row_off) = o;
            }
            EB_ARRIVE(3);
            EB_WAITM();
            // ---- E4: g_a2 = D4 * [h2 > 0] -> Ga2 ----
            {
                uint32_t v[16];
                tmem_ld16(tmem_lane + TM_A + 16u * (uint32_t)qtr, v);
                tmem_ld_wait();
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {
                    const uint4 hb = *reinterpret_cast<const uint4*>(sH2 + (uint32_t)(2 * qtr + qq) * 2048u + row_off);
                    const uint32_t hw[4] = {hb.x, hb.y, hb.z, hb.w};
                    uint32_t ow[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)        // relu outputs are >= 0: positive <=> non-zero (one HSET2 + LOP3 per pair)
                        ow[k] = pack_bf16(__uint_as_float(v[qq * 8 + 2 * k]), __uint_as_float(v[qq * 8 + 2 * k + 1])) & bf16x2_nz_mask(hw[k]);
                    *reinterpret_cast<uint4*>(sGa2 + (uint32_t)(2 * qtr + qq) * 2048u + row_off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
            }
            EB_ARRIVE(4);
            EB_WAITM();
            // ---- E5: g_a1 = D5 * [h1 > 0] -> Ga1 ----
            {
                uint32_t v[16];
                tmem_ld16(tmem_lane + TM_B + 16u * (uint32_t)qtr, v);
                tmem_ld_wait();
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {
                    const uint4 hb = *reinterpret_cast<const uint4*>(sH1 + (uint32_t)(2 * qtr + qq) * 2048u + row_off);
                    const uint32_t hw[4] = {hb.x, hb.y, hb.z, hb.w};
                    uint32_t ow[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ow[k] = pack_bf16(__uint_as_float(v[qq * 8 + 2 * k]), __uint_as_float(v[qq * 8 + 2 * k + 1])) & bf16x2_nz_mask(hw[k]);
                    *reinterpret_cast<uint4*>(sGa1 + (uint32_t)(2 * qtr + qq) * 2048u + row_off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
            }
            EB_ARRIVE(5);
            EB_WAITM();                                        // D6 and every weight-gradient MMA of the tile are complete
            // ---- E6: g_p (D6, k' = 4c + f) -> zero-padded fp32 planes [c][f][12][20]; pass-through g1 -> base plane ----
            {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int o = zoff[q];
                    if (o >= 0) {
                        if (o & (1 << 30)) {
                            float* pr = sPl + (o & ~(1 << 30));
                            *reinterpret_cast<float2*>(pr) = make_float2(0.f, 0.f);
                            *reinterpret_cast<float2*>(pr + 18) = make_float2(0.f, 0.f);
                        } else {
                            *reinterpret_cast<float4*>(sPl + o) = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
                // thread -> channels 6q .. 6q+5 = columns 24q .. 24q+23
                if (24 * qtr < g.K1) {
                    uint32_t v[24];
                    tmem_ld16(tmem_lane + TM_A + 24u * (uint32_t)qtr, v);
                    tmem_ld8(tmem_lane + TM_A + 24u * (uint32_t)qtr + 16u, v + 16);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const int c = 6 * qtr + i;
                        if (c < C) {
                            float* pc = sPl + (c * 3) * EB_PP + (py + 2) * EB_PS + px + 2;
                            pc[0] = __uint_as_float(v[4 * i + 0]);
                            pc[EB_PP] = __uint_as_float(v[4 * i + 1]);
                            pc[2 * EB_PP] = __uint_as_float(v[4 * i + 2]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (8 * qtr + i < C) sBase[((8 * qtr + i) * T2_TH + py) * T2_TW + px] = g1[i];
            }
            bar_sync_n(1, EB_NCOMP);
            // ---- transposed depthwise conv -> dL/dx_t, dL/dgoal.  item = (channel pair, 5-row block); lane = (column, channel) ----
            {
                const int hc = lane >> 4;
                for (int item = warp; item < 2 * g.npairs; item += 16) {
                    const int cp = item >> 1, vb = item & 1, c = 2 * cp + hc;
                    if (c < C) {
                        const float* P0 = sPl + (c * 3) * EB_PP;
                        const float* w = sWp + c * 27;
                        float* gob = a.g_out + ((size_t)b * C + c) * plane;
                        float* ggb = a.g_goal + ((size_t)b * C + c) * plane;
                        const int ox = (lane & 15) + 1;
                        float out[5];
                        eb_conv_t<5>(P0, w, 5 * vb, ox, out);
                        const int xx = x0 + ox - 1;
#pragma unroll
                        for (int k = 0; k < 5; ++k) {
                            const int oy = 5 * vb + k, yy = y0 - 1 + oy;
                            if (yy < 0 || yy >= H || xx >= W) continue;       // zero padding: nothing folds back
                            float vx = out[k];
                            if (oy >= 1 && oy <= T2_TH) vx += sBase[(c * T2_TH + oy - 1) * T2_TW + ox - 1];
                            atomicAdd(gob + (size_t)yy * W + xx, vx);
                            if (sPre[oy * 18 + ox] != 0.0f) atomicAdd(ggb + (size_t)yy * W + xx, out[k]);
                        }
                    }
                    if (lane < 20) {     // the two ring columns: 2 channels x 5 rows x 2 sides
                        const int hc2 = lane / 10, rem = lane % 10, c2 = 2 * cp + hc2;
                        if (c2 < C) {
                            const int oy = 5 * vb + (rem >> 1), ox = (rem & 1) ? T2_TW + 1 : 0;
                            float out[1];
                            eb_conv_t<1>(sPl + (c2 * 3) * EB_PP, sWp + c2 * 27, oy, ox, out);
                            const int yy = y0 - 1 + oy, xx = x0 - 1 + ox;
                            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                                atomicAdd(a.g_out + ((size_t)b * C + c2) * plane + (size_t)yy * W + xx, out[0]);
                                if (sPre[oy * 18 + ox] != 0.0f) atomicAdd(a.g_goal + ((size_t)b * C + c2) * plane + (size_t)yy * W + xx, out[0]);
                            }
                        }
                    }
                }
            }
            // ---- gwp[3c+f][a][b] += sum_cells g_p[3c+f][cell] * xin[c][cell + (a,b)]: warps 0 .. npairs-1, lane = (column, channel) ----
            if (warp < g.npairs) {
                const int hc = lane >> 4, pxx = lane & 15, c = 2 * warp + hc;
                if (c < C) {
#pragma unroll
                    for (int vb = 0; vb < 2; ++vb) {
                        const float* ch = sXin + (c * T2_XR + 4 * vb) * 18 + pxx;
                        float v[6][3];
#pragma unroll
                        for (int k = 0; k < 6; ++k)
#pragma unroll
                            for (int j = 0; j < 3; ++j) v[k][j] = ch[k * 18 + j];
#pragma unroll
                        for (int f = 0; f < 3; ++f) {
                            const float* P = sPl + (c * 3 + f) * EB_PP + (4 * vb + 2) * EB_PS + pxx + 2;
                            float gp[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) gp[k] = P[k * EB_PS];
#pragma unroll
                            for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                                for (int bb = 0; bb < 3; ++bb)
#pragma unroll
                                    for (int k = 0; k < 4; ++k) wpacc[f * 9 + aa * 3 + bb] = fmaf(gp[k], v[k + aa][bb], wpacc[f * 9 + aa * 3 + bb]);
                        }
                    }
                }
            }
            bar_sync_n(1, EB_NCOMP);     // the planes overlay the operand region the next tile writes
        }
        // ---- flush ----
        {
            const int j = (warp & 3) * 32 + lane;      // TMEM lane = hidden unit (rows of D7 .. D10); only j < 64 is real
            const int K = 3 * C;
            if ((warp & 3) < 2) {
                uint32_t v[32];
                if (qtr == 0) {          // D7: gwc[c][j]; D10: gbb[j]
                    tmem_ld32(tmem_lane + TM_D7, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (c < C) atomicAdd(a.gwc + c * 64 + j, __uint_as_float(v[c]));
                    tmem_ld16(tmem_lane + TM_D10, v);
                    tmem_ld_wait();
                    atomicAdd(a.gbb + j, __uint_as_float(v[0]));
                } else if (qtr == 1) {   // D8: gwb[j][k]
#pragma unroll 1
                    for (int k0 = 0; k0 < 64; k0 += 32) {
                        tmem_ld32(tmem_lane + TM_D8 + (uint32_t)k0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) atomicAdd(a.gwb + j * 64 + k0 + i, __uint_as_float(v[i]));
                    }
                } else if (qtr == 2) {   // D9: gwa[j][3c+f] (k' = 4c + f), gba[j] (k' = 8 * npairs)
#pragma unroll 1
                    for (int k0 = 0; k0 < g.K1; k0 += 32) {
                        tmem_ld32(tmem_lane + TM_D9 + (uint32_t)k0, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int kp = k0 + i;
                            if (kp >= g.K1) continue;
                            const int c = kp >> 2, f = kp & 3;
                            if (kp < 8 * g.npairs) { if (c < C && f < 3) atomicAdd(a.gwa + j * K + 3 * c + f, __uint_as_float(v[i])); }
                            else if (kp == 8 * g.npairs) atomicAdd(a.gba + j, __uint_as_float(v[i]));
                        }
                    }
                }
            }
            // gwp: reduce the per-thread accumulators over the 16 columns of a half-warp
            if (warp < g.npairs) {
                const int hc = lane >> 4, c = 2 * warp + hc;
#pragma unroll
                for (int i = 0; i < 27; ++i) {
                    float s = wpacc[i];
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if ((lane & 15) == 0 && c < C) atomicAdd(a.gwp + c * 27 + i, s);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem_base, 512u);
}

bool enc_tc_bwd_supported(const NcaEncDesc* d) {
    if (!enc_tc_supported(d)) return false;
    EncTcGeom g;
    etc_make_geom(d, &g);
    return g.K1 <= 96 && etb_smem(g).total <= 227u * 1024u;
}

int enc_tc_make_gmap(const NcaEncDesc* d, const float* gnext, EncTcMaps* m) {
    return t2_make_map((CUtensorMap*)m->x, gnext, 1, (size_t)d->B * d->C * d->H * d->W, d->B, d->C, d->H, d->W, T2_TH, T2_TW);
}

int enc_tc_backward_step(const NcaEncDesc* d, const NcaEncWeights* w, const void* ws, const EncTcMaps* m, const EncTcMaps* gm,
                         int slot_in, const uint8_t* life, float* g_out, float* g_goal, const NcaEncWeightGrads* gw,
                         const FireMask& fm, cudaStream_t s) {
    EncTcBwdArgs a;
    etc_make_geom(d, &a.g);
    a.clampv = d->clamp; a.slot_in = slot_in; a.life = life; a.g_out = g_out; a.g_goal = g_goal;
    a.Wa = (const __nv_bfloat16*)ws;
    a.Wb = a.Wa + (size_t)a.g.K1 * 64;
    a.Wc = a.Wb + 64 * 64;
    a.bb = w->bb; a.wp = w->wp;
    a.gwp = gw->wp; a.gwa = gw->wa; a.gba = gw->ba; a.gwb = gw->wb; a.gbb = gw->bb; a.gwc = gw->wc;
    a.fm = fm;
    a.tl = t2_make_tiles(d->B, d->H, d->W);
    const size_t smem = etb_smem(a.g).total;
    int grid = t2_num_sms();
    if (grid > a.tl.n_tiles) grid = a.tl.n_tiles;
    if (d->C == 20) {
        NCA_CUDA_OK(cudaFuncSetAttribute(enc_bwd_tc_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        enc_bwd_tc_kernel<20><<<grid, EB_NTHREADS, smem, s>>>(*(const CUtensorMap*)m->x, *(const CUtensorMap*)m->l, *(const CUtensorMap*)m->g,
                                                              *(const CUtensorMap*)gm->x, a);
    } else {
        NCA_CUDA_OK(cudaFuncSetAttribute(enc_bwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        enc_bwd_tc_kernel<0><<<grid, EB_NTHREADS, smem, s>>>(*(const CUtensorMap*)m->x, *(const CUtensorMap*)m->l, *(const CUtensorMap*)m->g,
                                                             *(const CUtensorMap*)gm->x, a);
    }
    NCA_LAUNCH_OK();
    return NCA_OK;
}
