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
    s.a1 = o; o += (uint32_t)(g.K1 / 8) * 2048u;      // A1, later A3 (first 16 KB)
    s.a2 = o; o += 8u * 2048u;
    s.total = o;
    return s;
}

template <int DUMMY>
__global__ void __launch_bounds__(ET2_NTHREADS) enc_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                   const __grid_constant__ CUtensorMap tm_l,
                                                                   const __grid_constant__ CUtensorMap tm_g, const EncTcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const EncTcGeom& g = a.g;
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
    uint8_t* sA3 = smem + L.a1;
    uint8_t* sA2 = smem + L.a2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = g.C, H = g.H, W = g.W;
    const size_t plane = (size_t)H * W;
    const int n_tiles = a.tl.n_tiles;
    const uint32_t stage_bytes = (uint32_t)(2 * C * T2_XR + ET2_LR) * T2_XS * 4u;

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
    const uint32_t TM_D1 = 0u, TM_D2 = 64u, TM_D3 = 0u;

    if (warp == 8) {
        // =========================== MMA / TMA warp ===========================
        const uint32_t id64 = umma_idesc_bf16(128, 64), id32 = umma_idesc_bf16(128, 32);
        const uint64_t dA1 = umma_desc(smem_u32(sA1), 2048u, 128u), dWa = umma_desc(smem_u32(sWa), 1024u, 128u);
        const uint64_t dA2 = umma_desc(smem_u32(sA2), 2048u, 128u), dWb = umma_desc(smem_u32(sWb), 1024u, 128u);
        const uint64_t dA3 = umma_desc(smem_u32(sA3), 2048u, 128u), dWc = umma_desc(smem_u32(sWc), 512u, 128u);
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
                    umma_ss(tmem_base + TM_D2, dA2 + (uint64_t)(ks * (4096 >> 4)), dWb + (uint64_t)(ks * (2048 >> 4)), id64, ks > 0);
                umma_commit(barM);
            }
            mbar_wait(barC, phC);
            phC ^= 1u;
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_ss(tmem_base + TM_D3, dA3 + (uint64_t)(ks * (4096 >> 4)), dWc + (uint64_t)(ks * (1024 >> 4)), id32, ks > 0);
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
                sPre[tid] = (g.liv < 0 || mx > g.thr) ? 1.0f : 0.0f;
            }
            // residual state of this thread's cell: channels 16*half .. 16*half+15
            float xres[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = 16 * half + i;
                xres[i] = c < C ? sX[(c * T2_XR + py + 1) * T2_XS + T2_XO + px + 1] : 0.0f;
            }
            bar_sync_n(1, 256);
            // ---- xin = x + goal * pre over the goal stage (nca.py:177) ----
            for (int i = tid; i < C * T2_XR * 18; i += 256) {
                const int q = i % 18, rr = (i / 18) % T2_XR, c = i / (18 * T2_XR);
                const int o = (c * T2_XR + rr) * T2_XS + T2_XO + q;
                sG[o] = fmaf(sG[o], sPre[rr * 18 + q], sX[o]);
            }
            bar_sync_n(1, 256);
            // ---- learned depthwise 3x3 -> A1: item = (channel pair, 4-row block); lane = (column, channel of the pair) ----
            {
                const int hc = lane >> 4, pxx = lane & 15;
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
            // ---- E1: h1 = relu(D1) -> A2 (ba rides in the bias chunk) ----
            {
                uint32_t v[32];
                tmem_ld32(tmem_lane + TM_D1 + 32u * (uint32_t)half, v);
                tmem_ld_wait();
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    uint4 o;
                    o.x = pack_bf16_relu(__uint_as_float(v[qq * 8 + 0]), __uint_as_float(v[qq * 8 + 1]));
                    o.y = pack_bf16_relu(__uint_as_float(v[qq * 8 + 2]), __uint_as_float(v[qq * 8 + 3]));
                    o.z = pack_bf16_relu(__uint_as_float(v[qq * 8 + 4]), __uint_as_float(v[qq * 8 + 5]));
                    o.w = pack_bf16_relu(__uint_as_float(v[qq * 8 + 6]), __uint_as_float(v[qq * 8 + 7]));
                    *reinterpret_cast<uint4*>(sA2 + (uint32_t)(4 * half + qq) * 2048u + row_off) = o;
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(barB);
            mbar_wait(barM, phM);
            phM ^= 1u;
            tc_fence_after();
            // ---- E2: h2 = relu(D2 + bb) -> A3 (over A1: MMA 1 is complete) ----
            {
                uint32_t v[32];
                tmem_ld32(tmem_lane + TM_D2 + 32u * (uint32_t)half, v);
                tmem_ld_wait();
                const float* bbp = sBb + 32 * half;
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    uint4 o;
                    o.x = pack_bf16_relu(__uint_as_float(v[qq * 8 + 0]) + bbp[qq * 8 + 0], __uint_as_float(v[qq * 8 + 1]) + bbp[qq * 8 + 1]);
                    o.y = pack_bf16_relu(__uint_as_float(v[qq * 8 + 2]) + bbp[qq * 8 + 2], __uint_as_float(v[qq * 8 + 3]) + bbp[qq * 8 + 3]);
                    o.z = pack_bf16_relu(__uint_as_float(v[qq * 8 + 4]) + bbp[qq * 8 + 4], __uint_as_float(v[qq * 8 + 5]) + bbp[qq * 8 + 5]);
                    o.w = pack_bf16_relu(__uint_as_float(v[qq * 8 + 6]) + bbp[qq * 8 + 6], __uint_as_float(v[qq * 8 + 7]) + bbp[qq * 8 + 7]);
                    *reinterpret_cast<uint4*>(sA3 + (uint32_t)(4 * half + qq) * 2048u + row_off) = o;
                }
            }
            fence_proxy_async();
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
                        const FireMask& fm, cudaStream_t s) {
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
    int occ = (int)((227 * 1024) / (smem + 1024));
    if (occ > 4) occ = 4;
    if (occ < 1) occ = 1;
    int grid = t2_num_sms() * occ;
    if (grid > a.tl.n_tiles) grid = a.tl.n_tiles;
    NCA_CUDA_OK(cudaFuncSetAttribute(enc_fwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    enc_fwd_tc_kernel<0><<<grid, ET2_NTHREADS, smem, s>>>(*(const CUtensorMap*)m->x, *(const CUtensorMap*)m->l, *(const CUtensorMap*)m->g, a);
    NCA_LAUNCH_OK();
    return NCA_OK;
}
