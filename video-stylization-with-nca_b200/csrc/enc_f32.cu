// ConditionedNCA step (EncoderConditioning/nca.py:152-209), fp32 CUDA-core path.
//
//   pre  = alive(x)                       3x3 max-pool (-inf padded) of the living channel > thr      (nca.py:152-163)
//   xin  = x + goal * pre                                                                             (nca.py:177)
//   p    = depthwise learned 3x3 conv of xin, zero padded, out channel j reads in channel j / 3       (nca.py:99-107)
//   out  = Wc relu(Wb relu(Wa p + ba) + bb)                                                           (nca.py:40-46)
//   x1   = x + fire * out ; post = alive(x1) ; x' = clamp(x1 * (pre & post), -clamp, clamp)           (nca.py:184-194)
//
// Forward = enc_step_kernel<false> (-> x1 in the workspace) + enc_life_kernel (post-update alive mask needs the
// updated 3x3 neighbourhood, so it is a second pass; it also records life = pre & post for the BPTT).
// BPTT    = enc_step_kernel<true>: recomputes the step from states[t], then walks the chain backwards.  One CTA owns a
// 4x32 tile; thread m < 128 owns cell m and evaluates the whole MLP chain for it with activations as [row][cell]
// matrices in shared memory and weights broadcast from shared memory; weight gradients are register accumulators
// spread over all 256 threads (outer products summed over the cells of every tile the CTA processes), flushed with
// one red.add per element at the end of the launch.
#include "nca_internal.h"

#define ET_TH 4
#define ET_TW 32
#define ET_TM 128
#define ET_TMS 132
#define ET_XR (ET_TH + 2)
#define ET_XS (ET_TW + 2)
#define ET_LR (ET_TH + 4)
#define ET_LS (ET_TW + 4)
#define ET_THREADS 256
#define ENC_HID 64
#define ENC_CP 24          // padded channel count of the output layer (C <= 21)

struct EncGeom {
    int B, C, H, W, liv;
    float thr, clampv;
};

// padded weight block in the workspace (floats)
#define ENC_OFF_WAT 0                                   // [64 k][64 j]  WaT[k][j] = wa[j][k]; row 3C = ba
#define ENC_OFF_WBT (ENC_OFF_WAT + 64 * 64)             // [64 j1][64 j2] = wb[j2][j1]
#define ENC_OFF_WCT (ENC_OFF_WBT + 64 * 64)             // [64 j][24 c]   = wc[c][j]
#define ENC_OFF_BB (ENC_OFF_WCT + 64 * ENC_CP)          // [64]
#define ENC_OFF_WP (ENC_OFF_BB + 64)                    // [64 j][9] (rows >= 3C zero), padded to 576
#define ENC_FWD_WFLOATS (ENC_OFF_WP + 576)
#define ENC_OFF_WA ENC_FWD_WFLOATS                      // [64 j][64 k]  = wa[j][k]
#define ENC_OFF_WB (ENC_OFF_WA + 64 * 64)               // [64 j2][64 j1] = wb[j2][j1]
#define ENC_OFF_WC (ENC_OFF_WB + 64 * 64)               // [24 c][64 j]  = wc[c][j]
#define ENC_BWD_WFLOATS (ENC_OFF_WC + ENC_CP * 64)

__global__ void enc_prep_weights_kernel(int C, const float* __restrict__ wp, const float* __restrict__ wa,
                                        const float* __restrict__ ba, const float* __restrict__ wb,
                                        const float* __restrict__ bb, const float* __restrict__ wc, float* __restrict__ ws) {
    const int K = 3 * C;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ENC_BWD_WFLOATS; i += gridDim.x * blockDim.x) {
        float v = 0.0f;
        if (i < ENC_OFF_WBT) { const int k = i / 64, j = i % 64; v = k < K ? wa[j * K + k] : (k == K ? ba[j] : 0.0f); }
        else if (i < ENC_OFF_WCT) { const int e = i - ENC_OFF_WBT, j1 = e / 64, j2 = e % 64; v = wb[j2 * 64 + j1]; }
        else if (i < ENC_OFF_BB) { const int e = i - ENC_OFF_WCT, j = e / ENC_CP, c = e % ENC_CP; v = c < C ? wc[c * 64 + j] : 0.0f; }
        else if (i < ENC_OFF_WP) { v = bb[i - ENC_OFF_BB]; }
        else if (i < ENC_FWD_WFLOATS) { const int e = i - ENC_OFF_WP; v = e < K * 9 ? wp[e] : 0.0f; }
        else if (i < ENC_OFF_WB) { const int e = i - ENC_OFF_WA, j = e / 64, k = e % 64; v = k < K ? wa[j * K + k] : 0.0f; }
        else if (i < ENC_OFF_WC) { v = wb[i - ENC_OFF_WB]; }
        else { const int e = i - ENC_OFF_WC, c = e / 64, j = e % 64; v = c < C ? wc[c * 64 + j] : 0.0f; }
        ws[i] = v;
    }
}

struct EncStepArgs {
    EncGeom g;
    const float* x_in;       // states[t]
    const float* goal;       // [B,C,H,W]
    const float* wsW;        // padded weights
    float* x1;               // forward: x + fire * out (all cells)
    // backward only
    const uint8_t* life;     // [B,H,W] life mask of this step
    const float* g_next;     // dL/d states[t+1] or NULL
    float* g_out;            // dL/d states[t], zeroed by the caller (red.add)
    float* g_goal;           // accumulated over steps (red.add)
    float* gwp; float* gwa; float* gba; float* gwb; float* gbb; float* gwc;   // reference layouts (red.add)
    FireMask fm;
    int tiles_x, tiles_y, n_tiles;
};

__device__ __forceinline__ float enc_fire(const FireMask& m, int b, int y, int x, int H, int W) {
    if (m.supplied) return m.supplied[((size_t)b * H + y) * W + x];
    return nca_fire(nca_philox_word((uint32_t)(y * W + x), (uint32_t)b, m.t, m.k0, m.k1), m.thr, 1);
}

// acc[n] += sum_{k < K} sIn[k][m] * sW[k][n]
template <int N>
__device__ __forceinline__ void enc_layer(const float* __restrict__ sIn, int K, const float* __restrict__ sW, int m, float acc[N]) {
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
        const float v = sIn[k * ET_TMS + m];
        const float4* w = reinterpret_cast<const float4*>(sW + k * N);
#pragma unroll
        for (int q = 0; q < N / 4; ++q) {
            const float4 ww = w[q];
            acc[4 * q + 0] = fmaf(v, ww.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(v, ww.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, ww.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(v, ww.w, acc[4 * q + 3]);
        }
    }
}

// acc[i][i2] += sum_m sA[ia + 16 i][m] * sB[ib + 16 i2][m]   (ia = tid >> 4, ib = tid & 15)
__device__ __forceinline__ void enc_wgrad_4x4(const float* __restrict__ sA, const float* __restrict__ sB, float acc[4][4]) {
    const float* ap = sA + (threadIdx.x >> 4) * ET_TMS;
    const float* bp = sB + (threadIdx.x & 15) * ET_TMS;
    for (int m = 0; m < ET_TM; m += 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = *reinterpret_cast<const float4*>(ap + 16 * i * ET_TMS + m);
            b[i] = *reinterpret_cast<const float4*>(bp + 16 * i * ET_TMS + m);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2)
                acc[i][i2] = fmaf(a[i].x, b[i2].x, fmaf(a[i].y, b[i2].y, fmaf(a[i].z, b[i2].z, fmaf(a[i].w, b[i2].w, acc[i][i2]))));
    }
}

static inline size_t enc_smem_floats(bool bwd) {
    size_t n = bwd ? ENC_BWD_WFLOATS : ENC_FWD_WFLOATS;
    n += (size_t)(64 + 64 + 64) * ET_TMS;                  // sP, sH1, sH2
    if (bwd) n += (size_t)ENC_CP * ET_TMS;                 // sGo
    n += (size_t)ENC_CP * ET_XR * ET_XS;                   // sXin
    n += ET_LR * ET_LS + ET_XR * ET_XS + 2 * ET_TM;        // sLiv, sPre, sFire, sLife
    return n;
}

template <bool BWD>
__global__ void __launch_bounds__(ET_THREADS, 1) enc_step_kernel(const EncStepArgs a) {
    extern __shared__ __align__(16) float smem[];
    const EncGeom& g = a.g;
    const int C = g.C, H = g.H, W = g.W, K = 3 * C;
    float* sW = smem;
    float* sP = sW + (BWD ? ENC_BWD_WFLOATS : ENC_FWD_WFLOATS);
    float* sH1 = sP + 64 * ET_TMS;
    float* sH2 = sH1 + 64 * ET_TMS;
    float* sGo = sH2 + 64 * ET_TMS;
    float* sXin = sGo + (BWD ? ENC_CP * ET_TMS : 0);
    float* sLiv = sXin + ENC_CP * ET_XR * ET_XS;
    float* sPre = sLiv + ET_LR * ET_LS;
    float* sFire = sPre + ET_XR * ET_XS;
    float* sLife = sFire + ET_TM;
    const float* sWaT = sW + ENC_OFF_WAT;
    const float* sWbT = sW + ENC_OFF_WBT;
    const float* sWcT = sW + ENC_OFF_WCT;
    const float* sBb = sW + ENC_OFF_BB;
    const float* sWp = sW + ENC_OFF_WP;
    const int tid = threadIdx.x;
    const size_t plane = (size_t)H * W;

    for (int i = tid; i < (BWD ? ENC_BWD_WFLOATS : ENC_FWD_WFLOATS) / 4; i += ET_THREADS)
        reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(a.wsW) + i);
    // rows of sP beyond the constant-1 row stay zero for the whole launch
    for (int i = tid; i < (64 - K - 1) * ET_TMS; i += ET_THREADS) sP[(K + 1) * ET_TMS + i] = 0.0f;

    // persistent weight-gradient accumulators (BWD)
    float accC[6], accB[4][4], accA[4][4], accBb = 0.0f, accP[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 6; ++i) accC[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { accB[i][j] = 0.0f; accA[i][j] = 0.0f; }

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        int tt = tile;
        const int x0 = (tt % a.tiles_x) * ET_TW; tt /= a.tiles_x;
        const int y0 = (tt % a.tiles_y) * ET_TH;
        const int b = tt / a.tiles_y;
        const float* xb = a.x_in + (size_t)b * C * plane;
        const float* gb = a.goal + (size_t)b * C * plane;
        // ---- stage: living channel (+2 ring, -inf outside), fire / life of the tile ----
        for (int i = tid; i < ET_LR * ET_LS; i += ET_THREADS) {
            const int yy = y0 - 2 + i / ET_LS, xx = x0 - 2 + i % ET_LS;
            float v = -INFINITY;
            if (g.liv >= 0 && yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(xb + g.liv * plane + (size_t)yy * W + xx);
            sLiv[i] = v;
        }
        if (tid < ET_TM) {
            const int gy = y0 + (tid >> 5), gx = x0 + (tid & 31);
            const bool in = gy < H && gx < W;
            sFire[tid] = in ? enc_fire(a.fm, b, gy, gx, H, W) : 0.0f;
            if (BWD) sLife[tid] = in ? (float)a.life[((size_t)b * H + gy) * W + gx] : 0.0f;
        }
        __syncthreads();
        for (int i = tid; i < ET_XR * ET_XS; i += ET_THREADS) {
            const int r = i / ET_XS, q = i % ET_XS;
            float mx = -INFINITY;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) mx = fmaxf(mx, sLiv[(r + dy) * ET_LS + q + dx]);
            sPre[i] = (g.liv < 0 || mx > g.thr) ? 1.0f : 0.0f;
        }
        __syncthreads();
        for (int i = tid; i < C * ET_XR * ET_XS; i += ET_THREADS) {
            const int q = i % ET_XS, r = (i / ET_XS) % ET_XR, c = i / (ET_XS * ET_XR);
            const int yy = y0 - 1 + r, xx = x0 - 1 + q;
            float v = 0.0f;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const size_t o = c * plane + (size_t)yy * W + xx;
                v = fmaf(__ldg(gb + o), sPre[r * ET_XS + q], __ldg(xb + o));
            }
            sXin[i] = v;
        }
        __syncthreads();
        // ---- perception -> sP rows j = 3c + f, constant-1 row at K ----
        for (int i = tid; i < (C + 1) * ET_TM; i += ET_THREADS) {
            const int m = i & (ET_TM - 1), c = i >> 7;
            const int py = m >> 5, px = m & 31;
            const bool in = (y0 + py < H) && (x0 + px < W);
            if (c == C) { sP[K * ET_TMS + m] = in ? 1.0f : 0.0f; continue; }
            float f0 = 0.f, f1 = 0.f, f2 = 0.f;
            if (in) {
                const float* xs = sXin + (c * ET_XR + py) * ET_XS + px;
                const float* w = sWp + c * 27;
#pragma unroll
                for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                    for (int bb = 0; bb < 3; ++bb) {
                        const float v = xs[aa * ET_XS + bb];
                        f0 = fmaf(w[aa * 3 + bb], v, f0);
                        f1 = fmaf(w[9 + aa * 3 + bb], v, f1);
                        f2 = fmaf(w[18 + aa * 3 + bb], v, f2);
                    }
            }
            sP[(3 * c + 0) * ET_TMS + m] = f0;
            sP[(3 * c + 1) * ET_TMS + m] = f1;
            sP[(3 * c + 2) * ET_TMS + m] = f2;
        }
        __syncthreads();
        // ---- forward chain for cell m = tid ----
        const int m = tid & (ET_TM - 1);
        const int py = m >> 5, px = m & 31;
        const int gy = y0 + py, gx = x0 + px;
        const bool inimg = gy < H && gx < W;
        const size_t pix = (size_t)gy * W + gx;
        if (tid < ET_TM) {
            float h[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) h[j] = 0.0f;
            enc_layer<64>(sP, K + 1, sWaT, m, h);
#pragma unroll
            for (int j = 0; j < 64; ++j) sH1[j * ET_TMS + m] = fmaxf(h[j], 0.0f);
#pragma unroll
            for (int j = 0; j < 64; ++j) h[j] = sBb[j];
            enc_layer<64>(sH1, 64, sWbT, m, h);
#pragma unroll
            for (int j = 0; j < 64; ++j) sH2[j * ET_TMS + m] = fmaxf(h[j], 0.0f);
            float o[ENC_CP];
#pragma unroll
            for (int c = 0; c < ENC_CP; ++c) o[c] = 0.0f;
            enc_layer<ENC_CP>(sH2, 64, sWcT, m, o);
            const float fire = sFire[m];
            if (!BWD) {
                if (inimg) {
#pragma unroll
                    for (int c = 0; c < ENC_CP; ++c)
                        if (c < C) {
                            const size_t off = ((size_t)b * C + c) * plane + pix;
                            a.x1[off] = fmaf(fire, o[c], __ldg(a.x_in + off));
                        }
                }
            } else {
                const float life = sLife[m];
#pragma unroll
                for (int c = 0; c < ENC_CP; ++c) {
                    float go = 0.0f;
                    if (c < C && inimg) {
                        const size_t off = ((size_t)b * C + c) * plane + pix;
                        const float x1 = fmaf(fire, o[c], __ldg(a.x_in + off));
                        const float v = x1 * life;
                        const float gn = a.g_next ? __ldg(a.g_next + off) : 0.0f;
                        const float g1 = (v >= -g.clampv && v <= g.clampv) ? gn * life : 0.0f;   // clamp then life mask
                        if (g1 != 0.0f) atomicAdd(a.g_out + off, g1);                            // residual path
                        go = fire * g1;
                    }
                    sGo[c * ET_TMS + m] = go;
                }
            }
        }
        if (BWD) {
            __syncthreads();
            // ---- gWc[c][j] += g_out (x) h2 ----
            {
                const int jb = tid & 63, ca = tid >> 6;
                const float* bp = sH2 + jb * ET_TMS;
                const float* ap = sGo + ca * 6 * ET_TMS;
                for (int mm = 0; mm < ET_TM; mm += 4) {
                    const float4 hv = *reinterpret_cast<const float4*>(bp + mm);
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const float4 gv = *reinterpret_cast<const float4*>(ap + i * ET_TMS + mm);
                        accC[i] = fmaf(gv.x, hv.x, fmaf(gv.y, hv.y, fmaf(gv.z, hv.z, fmaf(gv.w, hv.w, accC[i]))));
                    }
                }
            }
            __syncthreads();
            // ---- g_h2 = Wc^T g_out ; g_a2 = g_h2 * [h2 > 0] -> sH2 ----
            float ga[64];
            if (tid < ET_TM) {
#pragma unroll
                for (int j = 0; j < 64; ++j) ga[j] = 0.0f;
                enc_layer<64>(sGo, C, sW + ENC_OFF_WC, m, ga);
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    const float hv = sH2[j * ET_TMS + m];
                    sH2[j * ET_TMS + m] = hv > 0.0f ? ga[j] : 0.0f;
                }
            }
            __syncthreads();
            // ---- gWb[j2][j1] += g_a2 (x) h1 ; gbb[j2] += sum g_a2 ----
            enc_wgrad_4x4(sH2, sH1, accB);
            if (tid < 64) {
                const float* r = sH2 + tid * ET_TMS;
                float s = 0.0f;
                for (int mm = 0; mm < ET_TM; mm += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(r + mm);
                    s += (v.x + v.y) + (v.z + v.w);
                }
                accBb += s;
            }
            __syncthreads();
            // ---- g_h1 = Wb^T g_a2 ; g_a1 = g_h1 * [h1 > 0] -> sH1 ----
            if (tid < ET_TM) {
#pragma unroll
                for (int j = 0; j < 64; ++j) ga[j] = 0.0f;
                enc_layer<64>(sH2, 64, sW + ENC_OFF_WB, m, ga);
#pragma unroll
                for (int j = 0; j < 64; ++j) {
                    const float hv = sH1[j * ET_TMS + m];
                    sH1[j * ET_TMS + m] = hv > 0.0f ? ga[j] : 0.0f;
                }
            }
            __syncthreads();
            // ---- gWa[j][k] += g_a1 (x) p  (row K of p is the constant 1 -> gba) ----
            enc_wgrad_4x4(sH1, sP, accA);
            __syncthreads();
            // ---- g_p = Wa^T g_a1 -> sP ----
            if (tid < ET_TM) {
#pragma unroll
                for (int j = 0; j < 64; ++j) ga[j] = 0.0f;
                enc_layer<64>(sH1, 64, sW + ENC_OFF_WA, m, ga);
#pragma unroll
                for (int j = 0; j < 64; ++j) sP[j * ET_TMS + m] = ga[j];
            }
            __syncthreads();
            // ---- gwp[j][tap] += sum_cells g_p[j][cell] * xin[c][cell + tap] ----
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int item = tid + q * ET_THREADS;
                if (item < K * 9) {
                    const int j = item / 9, tap = item % 9, c = j / 3;
                    const float* gp = sP + j * ET_TMS;
                    const float* xs = sXin + (c * ET_XR + tap / 3) * ET_XS + tap % 3;
                    float s = 0.0f;
                    for (int r = 0; r < ET_TH; ++r)
#pragma unroll 8
                        for (int qx = 0; qx < ET_TW; ++qx) s = fmaf(gp[r * ET_TW + qx], xs[r * ET_XS + qx], s);
                    accP[q] += s;
                }
            }
            // ---- transposed depthwise conv -> g_x, g_goal ----
            float* gob = a.g_out + (size_t)b * C * plane;
            float* ggb = a.g_goal + (size_t)b * C * plane;
            for (int i = tid; i < C * ET_XR * ET_XS; i += ET_THREADS) {
                const int rx = i % ET_XS, ry = (i / ET_XS) % ET_XR, c = i / (ET_XS * ET_XR);
                const int yy = y0 - 1 + ry, xx = x0 - 1 + rx;
                if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                float v = 0.0f;
#pragma unroll
                for (int aa = 0; aa < 3; ++aa) {
                    const int cy = ry - aa;
                    if (cy < 0 || cy >= ET_TH) continue;
#pragma unroll
                    for (int bb = 0; bb < 3; ++bb) {
                        const int cx = rx - bb;
                        if (cx < 0 || cx >= ET_TW) continue;
                        const int cell = cy * ET_TW + cx;
                        const float* w = sWp + c * 27 + aa * 3 + bb;
                        v = fmaf(w[0], sP[(3 * c) * ET_TMS + cell], v);
                        v = fmaf(w[9], sP[(3 * c + 1) * ET_TMS + cell], v);
                        v = fmaf(w[18], sP[(3 * c + 2) * ET_TMS + cell], v);
                    }
                }
                if (v != 0.0f) {
                    const size_t o = c * plane + (size_t)yy * W + xx;
                    atomicAdd(gob + o, v);
                    if (sPre[ry * ET_XS + rx] != 0.0f) atomicAdd(ggb + o, v);
                }
            }
        }
        __syncthreads();
    }
    if (BWD) {
        // ---- flush the weight-gradient partial sums (reference layouts) ----
        {
            const int jb = tid & 63, ca = tid >> 6;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int c = ca * 6 + i;
                if (c < C) atomicAdd(a.gwc + c * 64 + jb, accC[i]);
            }
        }
        const int ia = tid >> 4, ib = tid & 15;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2) {
                const int r = ia + 16 * i, q = ib + 16 * i2;
                atomicAdd(a.gwb + r * 64 + q, accB[i][i2]);
                if (q < K) atomicAdd(a.gwa + r * K + q, accA[i][i2]);
                else if (q == K) atomicAdd(a.gba + r, accA[i][i2]);
            }
        if (tid < 64) atomicAdd(a.gbb + tid, accBb);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int item = tid + q * ET_THREADS;
            if (item < K * 9) atomicAdd(a.gwp + item, accP[q]);
        }
    }
}

// post-update life mask + clamp: x' = clamp(x1 * (alive(x) & alive(x1)), -clamp, clamp)  (nca.py:190-194)
__global__ void enc_life_kernel(EncGeom g, const float* __restrict__ x, const float* __restrict__ x1, float* __restrict__ xo,
                                uint8_t* __restrict__ life_out) {
    const int H = g.H, W = g.W, C = g.C;
    const size_t plane = (size_t)H * W, n = (size_t)g.B * plane;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // programmatic dependent launch (no-ops otherwise)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int xx = (int)(i % W), yy = (int)((i / W) % H);
        const size_t b = i / plane;
        float life = 1.0f;
        if (g.liv >= 0) {
            const float* l0 = x + (b * C + g.liv) * plane;
            const float* l1 = x1 + (b * C + g.liv) * plane;
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int y2 = yy + dy, x2 = xx + dx;
                    if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) {
                        m0 = fmaxf(m0, __ldg(l0 + (size_t)y2 * W + x2));
                        m1 = fmaxf(m1, __ldg(l1 + (size_t)y2 * W + x2));
                    }
                }
            life = (m0 > g.thr && m1 > g.thr) ? 1.0f : 0.0f;
        }
        if (life_out) life_out[i] = (uint8_t)life;
        const size_t o = b * C * plane + (size_t)yy * W + xx;
        for (int c = 0; c < C; ++c) {
            const float v = __ldg(x1 + o + c * plane) * life;
            xo[o + c * plane] = fminf(fmaxf(v, -g.clampv), g.clampv);
        }
    }
}

// ---- host side -------------------------------------------------------------------------------------------
static int enc_make_geom(const NcaEncDesc* d, EncGeom* g) {
    NCA_CHECK_ARG(d != nullptr, "desc is NULL");
    NCA_CHECK_ARG(d->B > 0 && d->C > 0 && d->H > 0 && d->W > 0, "B,C,H,W must be positive");
    NCA_CHECK_ARG(d->C <= 21, "ConditionedNCA: C=%d > 21 is not supported", d->C);
    NCA_CHECK_ARG(d->hid == ENC_HID, "ConditionedNCA: hidden width must be 64 (nca.py:40-46), got %d", d->hid);
    NCA_CHECK_ARG(d->living_dim < d->C, "living_dim=%d out of range", d->living_dim);
    NCA_CHECK_ARG(d->mask_mode == NCA_MASK_SUPPLIED || d->mask_mode == NCA_MASK_PHILOX, "bad mask_mode");
    NCA_CHECK_ARG(d->precision == NCA_PREC_FP32 || d->precision == NCA_PREC_BF16 || d->precision == NCA_PREC_F16X3, "bad precision %d", d->precision);
    NCA_CHECK_ARG((long long)d->H * d->W < (1ll << 31), "H*W too large");
    g->B = d->B; g->C = d->C; g->H = d->H; g->W = d->W; g->liv = d->living_dim < 0 ? -1 : d->living_dim;
    g->thr = d->alive_thr; g->clampv = d->clamp;
    return NCA_OK;
}
static int enc_check_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        nca_set_error("no CUDA device available (%s); libnca_b200 has no CPU fallback", cudaGetErrorString(e));
        cudaGetLastError();
        return NCA_ERR_CUDA;
    }
    return NCA_OK;
}
static int enc_num_sms() { return nca_sm_count(); }
static FireMask enc_mask(const NcaEncDesc* d, const EncGeom& g, const float* masks, uint64_t seed, int t0, int t) {
    FireMask m;
    m.supplied = d->mask_mode == NCA_MASK_SUPPLIED ? masks + (size_t)t * g.B * g.H * g.W : nullptr;
    m.k0 = (uint32_t)(seed & 0xffffffffu); m.k1 = (uint32_t)(seed >> 32);
    m.t = (uint32_t)(t0 + t);
    m.thr = nca_fire_threshold(d->fire_rate, 1);
    return m;
}

extern "C" {

size_t nca_enc_workspace_bytes(const NcaEncDesc* d, int32_t backward) {
    EncGeom g;
    if (enc_make_geom(d, &g)) return 0;
    const size_t n = nca_align_up((size_t)g.B * g.C * g.H * g.W, 64);
    // [padded fp32 weights | x1 (forward) or 2 state-gradient buffers (backward) | tcgen05 operand images (bf16)]
    return (nca_align_up(ENC_BWD_WFLOATS, 64) + (backward ? 2 * n : n)) * sizeof(float) + enc_tc_weight_bytes(d);
}

int nca_enc_forward(const NcaEncDesc* d, const NcaEncWeights* w, const float* goal, const float* masks, uint64_t seed,
                    int32_t t0, int32_t T, int32_t keep_history, float* states, uint8_t* life_hist, void* workspace,
                    size_t workspace_bytes, void* stream) {
    EncGeom g;
    int rc = enc_make_geom(d, &g);
    if (rc) return rc;
    NCA_CHECK_ARG(w && w->wp && w->wa && w->ba && w->wb && w->bb && w->wc, "weights are NULL");
    NCA_CHECK_ARG(goal && states && T >= 0, "goal / states is NULL or T < 0");
    NCA_CHECK_ARG(d->mask_mode != NCA_MASK_SUPPLIED || masks != nullptr, "mask_mode == SUPPLIED needs a masks pointer");
    NCA_CHECK_ARG(!keep_history || life_hist != nullptr, "keep_history needs a life_hist buffer");
    rc = enc_check_device();
    if (rc) return rc;
    if (workspace == nullptr || workspace_bytes < nca_enc_workspace_bytes(d, 0)) {
        nca_set_error("workspace too small: %zu < %zu", workspace_bytes, nca_enc_workspace_bytes(d, 0));
        return NCA_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    float* wsW = (float*)workspace;
    float* x1 = wsW + nca_align_up(ENC_BWD_WFLOATS, 64);
    const size_t n = (size_t)g.B * g.C * g.H * g.W, cells = (size_t)g.B * g.H * g.W;
    const int lgrid = (int)((cells + 255) / 256 < (size_t)enc_num_sms() * 8 ? (cells + 255) / 256 : (size_t)enc_num_sms() * 8);
    if (enc_tc_supported(d) && T > 0) {
        // tcgen05 path: x1 by enc_fwd_tc_kernel, then the life / clamp pass
        void* wsT = (void*)(x1 + nca_align_up(n, 64));
        rc = enc_tc_prep_weights(d, w, wsT, s);
        if (rc) return rc;
        EncTcMaps maps;
        rc = enc_tc_make_maps(d, states, keep_history ? T + 1 : 2, goal, &maps);
        if (rc) return rc;
        for (int t = 0; t < T; ++t) {
            const int si = keep_history ? t : (t & 1), so = keep_history ? t + 1 : ((t + 1) & 1);
            // both kernels of a step launch programmatically: each starts under its predecessor's tail and waits
            // (griddepcontrol.wait) before it touches the predecessor's output
            rc = enc_tc_forward_step(d, w, wsT, &maps, si, x1, enc_mask(d, g, masks, seed, t0, t), s, t > 0);
            if (rc) return rc;
            {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)lgrid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = s;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                NCA_CUDA_OK(cudaLaunchKernelEx(&cfg, enc_life_kernel, g, (const float*)(states + (size_t)si * n), (const float*)x1,
                                               states + (size_t)so * n, keep_history ? life_hist + (size_t)t * cells : (uint8_t*)nullptr));
            }
            NCA_LAUNCH_OK();
        }
        return NCA_OK;
    }
    enc_prep_weights_kernel<<<32, 256, 0, s>>>(g.C, w->wp, w->wa, w->ba, w->wb, w->bb, w->wc, wsW);
    NCA_LAUNCH_OK();
    const size_t smem = enc_smem_floats(false) * sizeof(float);
    NCA_CUDA_OK(cudaFuncSetAttribute(enc_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EncStepArgs a = {};
    a.g = g; a.goal = goal; a.wsW = wsW; a.x1 = x1;
    a.tiles_x = (g.W + ET_TW - 1) / ET_TW; a.tiles_y = (g.H + ET_TH - 1) / ET_TH; a.n_tiles = g.B * a.tiles_x * a.tiles_y;
    const int grid = a.n_tiles < enc_num_sms() ? a.n_tiles : enc_num_sms();
    for (int t = 0; t < T; ++t) {
        const float* xin = keep_history ? states + (size_t)t * n : states + (size_t)(t & 1) * n;
        float* xout = keep_history ? states + (size_t)(t + 1) * n : states + (size_t)((t + 1) & 1) * n;
        a.x_in = xin; a.fm = enc_mask(d, g, masks, seed, t0, t);
        enc_step_kernel<false><<<grid, ET_THREADS, smem, s>>>(a);
        NCA_LAUNCH_OK();
        enc_life_kernel<<<lgrid, 256, 0, s>>>(g, xin, x1, xout, keep_history ? life_hist + (size_t)t * cells : nullptr);
        NCA_LAUNCH_OK();
    }
    return NCA_OK;
}

int nca_enc_backward(const NcaEncDesc* d, const NcaEncWeights* w, const float* goal, const float* masks, uint64_t seed,
                     int32_t t0, int32_t T, const float* states, const uint8_t* life_hist, const float* g_final, float* gx0,
                     float* g_goal, const NcaEncWeightGrads* gw, void* workspace, size_t workspace_bytes, void* stream) {
    EncGeom g;
    int rc = enc_make_geom(d, &g);
    if (rc) return rc;
    NCA_CHECK_ARG(w && w->wp && w->wa && w->ba && w->wb && w->bb && w->wc, "weights are NULL");
    NCA_CHECK_ARG(gw && gw->wp && gw->wa && gw->ba && gw->wb && gw->bb && gw->wc, "weight-gradient outputs are NULL");
    NCA_CHECK_ARG(goal && states && life_hist && gx0 && g_goal && T >= 0, "goal / states / life_hist / gx0 / g_goal is NULL or T < 0");
    NCA_CHECK_ARG(d->mask_mode != NCA_MASK_SUPPLIED || masks != nullptr, "mask_mode == SUPPLIED needs a masks pointer");
    rc = enc_check_device();
    if (rc) return rc;
    if (workspace == nullptr || workspace_bytes < nca_enc_workspace_bytes(d, 1)) {
        nca_set_error("workspace too small: %zu < %zu", workspace_bytes, nca_enc_workspace_bytes(d, 1));
        return NCA_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)g.B * g.C * g.H * g.W, nb = n * sizeof(float), cells = (size_t)g.B * g.H * g.W;
    const int K = 3 * g.C;
    float* wsW = (float*)workspace;
    float* gbuf[2] = {wsW + nca_align_up(ENC_BWD_WFLOATS, 64), wsW + nca_align_up(ENC_BWD_WFLOATS, 64) + nca_align_up(n, 64)};
    enc_prep_weights_kernel<<<32, 256, 0, s>>>(g.C, w->wp, w->wa, w->ba, w->wb, w->bb, w->wc, wsW);
    NCA_LAUNCH_OK();
    NCA_CUDA_OK(cudaMemsetAsync(g_goal, 0, nb, s));
    NCA_CUDA_OK(cudaMemsetAsync(gw->wp, 0, (size_t)K * 9 * sizeof(float), s));
    NCA_CUDA_OK(cudaMemsetAsync(gw->wa, 0, (size_t)64 * K * sizeof(float), s));
    NCA_CUDA_OK(cudaMemsetAsync(gw->ba, 0, 64 * sizeof(float), s));
    NCA_CUDA_OK(cudaMemsetAsync(gw->wb, 0, 64 * 64 * sizeof(float), s));
    NCA_CUDA_OK(cudaMemsetAsync(gw->bb, 0, 64 * sizeof(float), s));
    NCA_CUDA_OK(cudaMemsetAsync(gw->wc, 0, (size_t)g.C * 64 * sizeof(float), s));
    if (T == 0) {
        if (g_final) NCA_CUDA_OK(cudaMemcpyAsync(gx0, g_final, nb, cudaMemcpyDeviceToDevice, s));
        else NCA_CUDA_OK(cudaMemsetAsync(gx0, 0, nb, s));
        return NCA_OK;
    }
    if (enc_tc_bwd_supported(d)) {
        // tcgen05 BPTT: one fused kernel per step; weight gradients go straight to the caller's buffers (zeroed above)
        void* wsT = (void*)(gbuf[1] + nca_align_up(n, 64));
        rc = enc_tc_prep_weights(d, w, wsT, s);
        if (rc) return rc;
        EncTcMaps maps, gm_final, gm[2];
        rc = enc_tc_make_maps(d, states, T + 1, goal, &maps);
        if (rc) return rc;
        for (int p = 0; p < 2; ++p) { rc = enc_tc_make_gmap(d, gbuf[p], &gm[p]); if (rc) return rc; }
        if (g_final) {
            NCA_CHECK_ARG((((uintptr_t)g_final) & 15u) == 0, "g_final must be 16-byte aligned");
            rc = enc_tc_make_gmap(d, g_final, &gm_final);
            if (rc) return rc;
        } else {
            NCA_CUDA_OK(cudaMemsetAsync(gbuf[T & 1], 0, nb, s));      // zeros stand in for dL/d states[T]
        }
        for (int t = T - 1; t >= 0; --t) {
            float* gout = t == 0 ? gx0 : gbuf[t & 1];
            NCA_CUDA_OK(cudaMemsetAsync(gout, 0, nb, s));
            const EncTcMaps* gmap = (t == T - 1) ? (g_final ? &gm_final : &gm[T & 1]) : &gm[(t + 1) & 1];
            rc = enc_tc_backward_step(d, w, wsT, &maps, gmap, t, life_hist + (size_t)t * cells, gout, g_goal, gw,
                                      enc_mask(d, g, masks, seed, t0, t), s);
            if (rc) return rc;
        }
        return NCA_OK;
    }
    const size_t smem = enc_smem_floats(true) * sizeof(float);
    if (smem > 227 * 1024) { nca_set_error("shared memory need %zu B exceeds 227 KB", smem); return NCA_ERR_UNSUPPORTED; }
    NCA_CUDA_OK(cudaFuncSetAttribute(enc_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    EncStepArgs a = {};
    a.g = g; a.goal = goal; a.wsW = wsW; a.g_goal = g_goal;
    a.gwp = gw->wp; a.gwa = gw->wa; a.gba = gw->ba; a.gwb = gw->wb; a.gbb = gw->bb; a.gwc = gw->wc;
    a.tiles_x = (g.W + ET_TW - 1) / ET_TW; a.tiles_y = (g.H + ET_TH - 1) / ET_TH; a.n_tiles = g.B * a.tiles_x * a.tiles_y;
    const int grid = a.n_tiles < enc_num_sms() ? a.n_tiles : enc_num_sms();
    const float* gnext = g_final;
    for (int t = T - 1; t >= 0; --t) {
        float* gout = t == 0 ? gx0 : gbuf[t & 1];
        NCA_CUDA_OK(cudaMemsetAsync(gout, 0, nb, s));
        a.x_in = states + (size_t)t * n; a.life = life_hist + (size_t)t * cells; a.g_next = gnext; a.g_out = gout;
        a.fm = enc_mask(d, g, masks, seed, t0, t);
        enc_step_kernel<true><<<grid, ET_THREADS, smem, s>>>(a);
        NCA_LAUNCH_OK();
        gnext = gout;
    }
    return NCA_OK;
}

}  // extern "C"
