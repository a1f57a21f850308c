// DyNCA step, fp32 CUDA-core path (NCA_PREC_FP32): forward step, BPTT step, perceive-only, weight packing.
// Reference semantics: ExtraChannels/models/dynca.py:71-128, ConditioneDyNCA/models/dynca.py:117-138.
//
// Both step kernels are persistent (grid = #SMs x occupancy, tiles strided over CTAs) so the padded weight
// matrices are loaded into shared memory once per launch and, in the backward kernel, the weight-gradient
// accumulators stay in registers across all tiles of a CTA and are flushed with one red.add per element.
#include "dynca_tile.cuh"
#include "nca_internal.h"

// ------------------------------------------------------------------------------------------------
// weight packing / unpacking
// ------------------------------------------------------------------------------------------------
__global__ void dynca_prep_weights_kernel(DyncaGeom g, const float* __restrict__ w1, const float* __restrict__ b1,
                                          const float* __restrict__ w2, const float* __restrict__ b2,
                                          float* __restrict__ W1p, float* __restrict__ W2p, float* __restrict__ W2t,
                                          float* __restrict__ b2p) {
    const int n1 = g.Ppad * g.FCpad, n2 = g.FCpad * g.CP;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + 2 * n2 + g.CP; i += gridDim.x * blockDim.x) {
        if (i < n1) {
            int k = i / g.FCpad, j = i % g.FCpad;
            float v = 0.0f;
            if (j < g.fc) { if (k < g.P) v = w1[j * g.P + k]; else if (k == g.P) v = b1[j]; }
            W1p[i] = v;
        } else if (i < n1 + n2) {
            int e = i - n1, j = e / g.CP, c = e % g.CP;
            W2p[e] = (j < g.fc && c < g.C) ? w2[c * g.fc + j] : 0.0f;
        } else if (i < n1 + 2 * n2) {
            int e = i - n1 - n2, c = e / g.FCpad, j = e % g.FCpad;
            W2t[e] = (j < g.fc && c < g.C) ? w2[c * g.fc + j] : 0.0f;
        } else {
            int c = i - n1 - 2 * n2;
            b2p[c] = c < g.C ? b2[c] : 0.0f;
        }
    }
}

__global__ void dynca_unpack_grads_kernel(DyncaGeom g, const float* __restrict__ gW1p, const float* __restrict__ gW2p,
                                          const float* __restrict__ gb2p, float* __restrict__ gw1, float* __restrict__ gb1,
                                          float* __restrict__ gw2, float* __restrict__ gb2) {
    const int n1 = g.fc * g.P, n2 = g.C * g.fc;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + g.fc + n2 + g.C; i += gridDim.x * blockDim.x) {
        if (i < n1) { int j = i / g.P, k = i % g.P; gw1[i] = gW1p[k * g.FCpad + j]; }
        else if (i < n1 + g.fc) { int j = i - n1; gb1[j] = gW1p[g.P * g.FCpad + j]; }
        else if (i < n1 + g.fc + n2) { int e = i - n1 - g.fc, c = e / g.fc, j = e % g.fc; gw2[e] = gW2p[j * g.CP + c]; }
        else { int c = i - n1 - g.fc - n2; gb2[c] = gb2p[c]; }
    }
}

// ------------------------------------------------------------------------------------------------
// forward step
// ------------------------------------------------------------------------------------------------
struct DyncaFwdArgs {
    DyncaGeom g;
    const float* x_in; float* x_out; const float* cond;
    const float* W1p; const float* W2p; const float* b2p;
    FireMask fm;
    int tiles_x, tiles_y, n_tiles;
};

static inline size_t dynca_fwd_smem_floats(const DyncaGeom& g) {
    size_t u1 = (size_t)dynca_stage_floats(g) + (size_t)g.Ppad * DT_TMS;
    size_t u2 = (size_t)g.FCpad * DT_TMS;
    return (size_t)g.Ppad * g.FCpad + (size_t)g.FCpad * g.CP + g.CP + DT_TM + (u1 > u2 ? u1 : u2);
}

template <int NS>
__global__ void __launch_bounds__(DT_THREADS) dynca_fwd_f32_kernel(const DyncaFwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    const DyncaGeom& g = a.g;
    float* sW1 = smem;
    float* sW2 = sW1 + g.Ppad * g.FCpad;
    float* sB2 = sW2 + g.FCpad * g.CP;
    float* sMask = sB2 + g.CP;
    float* sU = sMask + DT_TM;
    float* sStage = sU;
    float* sZ = sU + dynca_stage_floats(g);
    float* sH = sU;   // overlays stage + Z once GEMM1 is done
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < g.Ppad * g.FCpad / 4; i += DT_THREADS)
        reinterpret_cast<float4*>(sW1)[i] = __ldg(reinterpret_cast<const float4*>(a.W1p) + i);
    for (int i = tid; i < g.FCpad * g.CP / 4; i += DT_THREADS)
        reinterpret_cast<float4*>(sW2)[i] = __ldg(reinterpret_cast<const float4*>(a.W2p) + i);
    if (tid < g.CP) sB2[tid] = a.b2p[tid];
    const size_t plane = (size_t)g.H * g.W;

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const DyncaTile t = dynca_tile_of(tile, a.tiles_x, a.tiles_y);
        if (tid < DT_TM) {
            int gy = t.y0 + (tid >> 5), gx = t.x0 + (tid & 31);
            sMask[tid] = (gy < g.H && gx < g.W) ? dynca_fire(a.fm, t.b, gy, gx, g.H, g.W) : 0.0f;
        }
        dynca_perceive_tile<NS>(g, a.x_in, a.cond, t, sStage, sZ);   // ends with a barrier (also covers sW*, sMask)
        float acc[8][8];
        dynca_gemm1(g, sZ, sW1, acc);
        __syncthreads();                                             // everyone done reading sZ
        {
            const int tx = tid & 15, ty = tid >> 4;
            if (ty * 8 < g.FCpad) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float* row = sH + (ty * 8 + j) * DT_TMS + tx * 4;
                    *reinterpret_cast<float4*>(row) = make_float4(fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f),
                                                                  fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
                    *reinterpret_cast<float4*>(row + 64) = make_float4(fmaxf(acc[4][j], 0.f), fmaxf(acc[5][j], 0.f),
                                                                       fmaxf(acc[6][j], 0.f), fmaxf(acc[7][j], 0.f));
                }
            }
        }
        __syncthreads();
        // GEMM2 + epilogue: lane -> 4 consecutive cells, warp -> 2 channels
        {
            float y[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int q = 0; q < 4; ++q) y[i][q] = 0.0f;
            const float* hp = sH + lane * 4;
            const float* wp = sW2 + warp * 2;
#pragma unroll 4
            for (int j = 0; j < g.fc; ++j) {
                float4 h = *reinterpret_cast<const float4*>(hp + j * DT_TMS);
                float2 w = *reinterpret_cast<const float2*>(wp + j * g.CP);
                y[0][0] = fmaf(h.x, w.x, y[0][0]); y[0][1] = fmaf(h.y, w.x, y[0][1]);
                y[0][2] = fmaf(h.z, w.x, y[0][2]); y[0][3] = fmaf(h.w, w.x, y[0][3]);
                y[1][0] = fmaf(h.x, w.y, y[1][0]); y[1][1] = fmaf(h.y, w.y, y[1][1]);
                y[1][2] = fmaf(h.z, w.y, y[1][2]); y[1][3] = fmaf(h.w, w.y, y[1][3]);
            }
            const int gy = t.y0 + (lane >> 3), gx0 = t.x0 + (lane & 7) * 4;
            if (gy < g.H) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int c = warp * 2 + i;
                    if (c >= g.C) continue;
                    const size_t off = ((size_t)t.b * g.C + c) * plane + (size_t)gy * g.W + gx0;
                    const float bias = sB2[c];
                    if ((g.W & 3) == 0 && gx0 + 3 < g.W) {
                        float4 xi = __ldg(reinterpret_cast<const float4*>(a.x_in + off));
                        float4 mk = *reinterpret_cast<const float4*>(sMask + lane * 4);
                        float4 o;
                        o.x = xi.x + (y[i][0] + bias) * mk.x; o.y = xi.y + (y[i][1] + bias) * mk.y;
                        o.z = xi.z + (y[i][2] + bias) * mk.z; o.w = xi.w + (y[i][3] + bias) * mk.w;
                        *reinterpret_cast<float4*>(a.x_out + off) = o;
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (gx0 + q < g.W) a.x_out[off + q] = __ldg(a.x_in + off + q) + (y[i][q] + bias) * sMask[lane * 4 + q];
                    }
                }
            }
        }
        __syncthreads();   // sH / sMask are rewritten by the next tile
    }
}

// ------------------------------------------------------------------------------------------------
// perceive only (DyNCA.perceive_multiscale as an op; tests + return_perception=True)
// ------------------------------------------------------------------------------------------------
struct DyncaPerceiveArgs {
    DyncaGeom g;
    const float* x; const float* cond; float* z;
    int tiles_x, tiles_y, n_tiles;
};
template <int NS>
__global__ void __launch_bounds__(DT_THREADS) dynca_perceive_kernel(const DyncaPerceiveArgs a) {
    extern __shared__ __align__(16) float smem[];
    const DyncaGeom& g = a.g;
    float* sStage = smem;
    float* sZ = smem + dynca_stage_floats(g);
    const size_t plane = (size_t)g.H * g.W;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const DyncaTile t = dynca_tile_of(tile, a.tiles_x, a.tiles_y);
        dynca_perceive_tile<NS>(g, a.x, a.cond, t, sStage, sZ);
        for (int i = threadIdx.x; i < g.P * DT_TM; i += DT_THREADS) {
            int k = i >> 7, m = i & 127;
            int gy = t.y0 + (m >> 5), gx = t.x0 + (m & 31);
            if (gy < g.H && gx < g.W) a.z[((size_t)t.b * g.P + k) * plane + (size_t)gy * g.W + gx] = sZ[k * DT_TMS + m];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// backward (BPTT) step:  given x_t and g_{t+1} = dL/dx_{t+1}  ->  g_t (atomically accumulated into a
// zeroed buffer) and per-CTA weight-gradient partial sums.
//   g_y = mask * g_{t+1};  gW2 += g_y (x) h;  g_h = W2^T g_y;  g_a = g_h * [a>0];  gW1 += g_a (x) z;
//   g_z = W1^T g_a;  g_t = g_{t+1} + Perceive^T(g_z)
// ------------------------------------------------------------------------------------------------
struct DyncaBwdArgs {
    DyncaGeom g;
    const float* x_in;          // states[t]
    const float* g_next;        // dL/d states[t+1] accumulated so far (may be NULL = zeros)
    const float* g_tap;         // [B,tap_c,H,W] extra gradient on tap_scale * states[t+1][:, :tap_c], or NULL
    int tap_c; float tap_scale;
    float* g_out;               // dL/d states[t]; zeroed by the caller, red.add target
    const float* cond;
    const float* W1p; const float* W2t;
    float* gW1p; float* gW2p; float* gb2p;   // global accumulators (red.add)
    FireMask fm;
    int tiles_x, tiles_y, n_tiles;
};

static inline size_t dynca_bwd_smem_floats(const DyncaGeom& g) {
    size_t stage = (size_t)dynca_stage_floats(g);
    size_t gcp = g.ns == 2 ? (size_t)4 * g.C * DT_PCH * DT_PCW : 0;
    size_t h = (size_t)g.FCpad * DT_TMS;
    size_t u = stage > h ? stage : h;
    if (gcp > u) u = gcp;
    return (size_t)g.Ppad * g.FCpad + (size_t)g.CP * g.FCpad + DT_TM + (size_t)g.CP * DT_TMS + (size_t)g.Ppad * DT_TMS + u;
}

template <int NS>
__global__ void __launch_bounds__(DT_THREADS, 1) dynca_bwd_f32_kernel(const DyncaBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    const DyncaGeom& g = a.g;
    float* sW1 = smem;                          // [Ppad][FCpad]
    float* sW2t = sW1 + g.Ppad * g.FCpad;       // [CP][FCpad]
    float* sMask = sW2t + g.CP * g.FCpad;       // [TM]
    float* sGy = sMask + DT_TM;                 // [CP][TMS]
    float* sZ = sGy + g.CP * DT_TMS;            // [Ppad][TMS]   z, later g_z
    float* sU = sZ + g.Ppad * DT_TMS;           // stage | h / g_a | coarse g_percept
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = g.C, H = g.H, W = g.W, FCpad = g.FCpad;
    const size_t plane = (size_t)H * W;

    for (int i = tid; i < g.Ppad * FCpad / 4; i += DT_THREADS)
        reinterpret_cast<float4*>(sW1)[i] = __ldg(reinterpret_cast<const float4*>(a.W1p) + i);
    for (int i = tid; i < g.CP * FCpad / 4; i += DT_THREADS)
        reinterpret_cast<float4*>(sW2t)[i] = __ldg(reinterpret_cast<const float4*>(a.W2t) + i);

    // persistent weight-gradient accumulators
    //   gW1p[k][j]: k = kx + 16*i, j = jy + 16*i2   (kx = warp*2 + lane/16, jy = lane%16)
    //   gW2p[j][c]: j = tid & 127, c = (tid>>7)*8 + i
    float w1acc[8][8];
    float w2acc[8];
    float b2acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        w2acc[i] = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) w1acc[i][j] = 0.0f;
    }
    const int kx = warp * 2 + (lane >> 4), jy = lane & 15;
    const int kEnd = g.P + 1;

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const DyncaTile t = dynca_tile_of(tile, a.tiles_x, a.tiles_y);
        if (tid < DT_TM) {
            int gy = t.y0 + (tid >> 5), gx = t.x0 + (tid & 31);
            sMask[tid] = (gy < H && gx < W) ? dynca_fire(a.fm, t.b, gy, gx, H, W) : 0.0f;
        }
        dynca_perceive_tile<NS>(g, a.x_in, a.cond, t, sU, sZ);
        // g_y = mask * g_{t+1}
        for (int i = tid; i < g.CP * DT_TM; i += DT_THREADS) {
            int c = i >> 7, m = i & 127;
            int gy = t.y0 + (m >> 5), gx = t.x0 + (m & 31);
            float v = 0.0f;
            if (c < C && gy < H && gx < W) v = sMask[m] * dynca_gnext(a.g_next, a.g_tap, a.tap_c, a.tap_scale, C, t.b, c, (size_t)gy * W + gx, plane);
            sGy[c * DT_TMS + m] = v;
        }
        float acc[8][8];
        dynca_gemm1(g, sZ, sW1, acc);
        __syncthreads();   // sGy complete; stage area (sU) free
        const int tx = tid & 15, ty = tid >> 4;
        const bool act = ty * 8 < FCpad;
        unsigned long long pos = 0ull;   // bit i*8+j: a > 0
        if (act) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int i = 0; i < 8; ++i) if (acc[i][j] > 0.0f) pos |= 1ull << (i * 8 + j);
                float* row = sU + (ty * 8 + j) * DT_TMS + tx * 4;
                *reinterpret_cast<float4*>(row) = make_float4(fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f),
                                                              fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
                *reinterpret_cast<float4*>(row + 64) = make_float4(fmaxf(acc[4][j], 0.f), fmaxf(acc[5][j], 0.f),
                                                                   fmaxf(acc[6][j], 0.f), fmaxf(acc[7][j], 0.f));
            }
            // g_h = W2^T g_y  (into acc), then g_a = g_h * [a > 0]
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
            for (int c = 0; c < C; ++c) {
                float4 ya = *reinterpret_cast<const float4*>(sGy + c * DT_TMS + tx * 4);
                float4 yb = *reinterpret_cast<const float4*>(sGy + c * DT_TMS + 64 + tx * 4);
                float4 wa = *reinterpret_cast<const float4*>(sW2t + c * FCpad + ty * 8);
                float4 wb = *reinterpret_cast<const float4*>(sW2t + c * FCpad + ty * 8 + 4);
                const float yv[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
                const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(yv[i], wv[j], acc[i][j]);
            }
        }
        __syncthreads();   // h complete in sU
        // gW2p[j][c] += sum_m h[j][m] g_y[c][m]
        {
            const int j = tid & 127, c0 = (tid >> 7) * 8;
            if (j < FCpad) {
                const float* hp = sU + j * DT_TMS;
                const float* yp = sGy + c0 * DT_TMS;
                for (int m = 0; m < DT_TM; m += 4) {
                    float4 h = *reinterpret_cast<const float4*>(hp + m);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4 y = *reinterpret_cast<const float4*>(yp + i * DT_TMS + m);
                        w2acc[i] = fmaf(h.x, y.x, fmaf(h.y, y.y, fmaf(h.z, y.z, fmaf(h.w, y.w, w2acc[i]))));
                    }
                }
            }
            if (tid < g.CP) {   // gb2[c] += sum_m g_y[c][m]
                const float* yp = sGy + tid * DT_TMS;
                float s = 0.0f;
                for (int m = 0; m < DT_TM; m += 4) {
                    float4 y = *reinterpret_cast<const float4*>(yp + m);
                    s += (y.x + y.y) + (y.z + y.w);
                }
                b2acc += s;
            }
        }
        __syncthreads();   // everyone done reading h
        if (act) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = ((pos >> (i * 8 + j)) & 1ull) ? acc[i][j] : 0.0f;
                float* row = sU + (ty * 8 + j) * DT_TMS + tx * 4;
                *reinterpret_cast<float4*>(row) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(row + 64) = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
        __syncthreads();   // g_a complete in sU
        // gW1p[k][j] += sum_m z[k][m] g_a[j][m]
        {
            const float* zp = sZ + kx * DT_TMS;
            const float* ap = sU + jy * DT_TMS;
            for (int m = 0; m < DT_TM; m += 4) {
                float4 av[8];
#pragma unroll
                for (int i2 = 0; i2 < 8; ++i2)
                    av[i2] = (jy + 16 * i2 < FCpad) ? *reinterpret_cast<const float4*>(ap + (16 * i2) * DT_TMS + m)
                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (16 * i < kEnd) {   // CTA-uniform
                        float4 z = (kx + 16 * i < kEnd) ? *reinterpret_cast<const float4*>(zp + (16 * i) * DT_TMS + m)
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int i2 = 0; i2 < 8; ++i2)
                            w1acc[i][i2] = fmaf(z.x, av[i2].x, fmaf(z.y, av[i2].y, fmaf(z.z, av[i2].z, fmaf(z.w, av[i2].w, w1acc[i][i2]))));
                    }
                }
            }
        }
        // g_z[k][m] = sum_j W1p[k][j] g_a[j][m], k < 4C : thread -> rows ty*4..+3, cells tx*4.. | 64+tx*4..
        float gz[4][8];
        {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int i = 0; i < 8; ++i) gz[kk][i] = 0.0f;
            if (ty * 4 < 4 * C) {
                const float* ap = sU + tx * 4;
                const float* wp = sW1 + (ty * 4) * FCpad;
                for (int j = 0; j < FCpad; j += 4) {
                    float4 w[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4*>(wp + kk * FCpad + j);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        float4 ga = *reinterpret_cast<const float4*>(ap + (j + jj) * DT_TMS);
                        float4 gb = *reinterpret_cast<const float4*>(ap + (j + jj) * DT_TMS + 64);
                        const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const float wv = jj == 0 ? w[kk].x : (jj == 1 ? w[kk].y : (jj == 2 ? w[kk].z : w[kk].w));
#pragma unroll
                            for (int i = 0; i < 8; ++i) gz[kk][i] = fmaf(wv, gv[i], gz[kk][i]);
                        }
                    }
                }
            }
        }
        __syncthreads();   // all reads of sZ (z) and sU (g_a) done
        if (ty * 4 < 4 * C) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float* row = sZ + (ty * 4 + kk) * DT_TMS + tx * 4;
                *reinterpret_cast<float4*>(row) = make_float4(gz[kk][0] * g.s0, gz[kk][1] * g.s0, gz[kk][2] * g.s0, gz[kk][3] * g.s0);
                *reinterpret_cast<float4*>(row + 64) = make_float4(gz[kk][4] * g.s0, gz[kk][5] * g.s0, gz[kk][6] * g.s0, gz[kk][7] * g.s0);
            }
        }
        __syncthreads();   // sZ now holds s0 * g_z (rows < 4C)
        // transposed perception of s0*g_z (+ residual pass-through of g_{t+1}) -> red.add into g_out
        dynca_scatter_tile<NS, DT_THREADS, false>(g, t, sZ, sU, a.g_out, a.g_next, a.g_tap, a.tap_c, a.tap_scale);
        __syncthreads();   // smem reused by the next tile
    }
    // ---- flush weight-gradient partial sums ----
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = kx + 16 * i;
        if (k >= kEnd) continue;
#pragma unroll
        for (int i2 = 0; i2 < 8; ++i2) {
            const int j = jy + 16 * i2;
            if (j < g.fc) atomicAdd(a.gW1p + k * FCpad + j, w1acc[i][i2]);
        }
    }
    {
        const int j = tid & 127, c0 = (tid >> 7) * 8;
        if (j < g.fc)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (c0 + i < C) atomicAdd(a.gW2p + j * g.CP + c0 + i, w2acc[i]);
        if (tid < C) atomicAdd(a.gb2p + tid, b2acc);
    }
}

// ------------------------------------------------------------------------------------------------
// EdgeExtractor (ConditioneDyNCA/models/dynca.py:182-213): zero-padded sobel_x / sobel_y / laplacian of a
// one-channel image, optional tanh
// ------------------------------------------------------------------------------------------------
__global__ void nca_edge_extract_kernel(int B, int H, int W, const float* __restrict__ img, int do_tanh, float* __restrict__ out) {
    const size_t plane = (size_t)H * W, n = (size_t)B * plane;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)((i / W) % H);
        const size_t b = i / plane;
        const float* p = img + b * plane;
        float v[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                int yy = y + a - 1, xx = x + c - 1;
                v[a][c] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(p + (size_t)yy * W + xx) : 0.0f;
            }
        float sx, sy, lap;
        dynca_filters(v, sx, sy, lap);
        if (do_tanh) { sx = tanhf(sx); sy = tanhf(sy); lap = tanhf(lap); }
        float* o = out + b * 3 * plane + (size_t)y * W + x;
        o[0] = sx; o[plane] = sy; o[2 * plane] = lap;
    }
}

__global__ void nca_philox_mask_kernel(int B, int H, int W, unsigned long long thr, int enc, uint32_t k0, uint32_t k1,
                                       int t0, const uint32_t* __restrict__ t0_dev, int T, float* __restrict__ out) {
    const size_t plane = (size_t)H * W, n = (size_t)T * B * plane;
    if (t0_dev) t0 += (int)__ldg(t0_dev);      // step counter kept on the device (CUDA-graph replays advance it)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t p = (uint32_t)(i % plane);
        const uint32_t b = (uint32_t)((i / plane) % B);
        const uint32_t t = (uint32_t)(i / (plane * B));
        out[i] = nca_fire(nca_philox_word(p, b, (uint32_t)t0 + t, k0, k1), thr, enc);
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static int nca_num_sms() { return nca_sm_count(); }

template <typename K>
static int dynca_set_smem(K kernel, size_t bytes) {
    if (bytes > 227 * 1024) { nca_set_error("shared memory need %zu B exceeds 227 KB", bytes); return NCA_ERR_UNSUPPORTED; }
    NCA_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return NCA_OK;
}

static void dynca_tiles(const DyncaGeom& g, int& tx, int& ty, int& n) {
    tx = (g.W + DT_TW - 1) / DT_TW; ty = (g.H + DT_TH - 1) / DT_TH; n = g.B * tx * ty;
}

size_t dynca_f32_weight_floats(const DyncaGeom& g) {
    return nca_align_up((size_t)g.Ppad * g.FCpad + 2 * (size_t)g.FCpad * g.CP + g.CP, 64);
}
size_t dynca_f32_grad_floats(const DyncaGeom& g) {
    return nca_align_up((size_t)g.Ppad * g.FCpad + (size_t)g.FCpad * g.CP + g.CP, 64);
}

int dynca_f32_prep_weights(const DyncaGeom& g, const NcaDyncaWeights* w, float* ws, cudaStream_t s) {
    float* W1p = ws; float* W2p = W1p + (size_t)g.Ppad * g.FCpad; float* W2t = W2p + (size_t)g.FCpad * g.CP;
    float* b2p = W2t + (size_t)g.FCpad * g.CP;
    dynca_prep_weights_kernel<<<32, 256, 0, s>>>(g, w->w1, w->b1, w->w2, w->b2, W1p, W2p, W2t, b2p);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int dynca_f32_forward_step(const DyncaGeom& g, const float* wsW, const float* x_in, float* x_out, const float* cond,
                           const FireMask& fm, cudaStream_t s) {
    DyncaFwdArgs a;
    a.g = g; a.x_in = x_in; a.x_out = x_out; a.cond = cond;
    a.W1p = wsW; a.W2p = a.W1p + (size_t)g.Ppad * g.FCpad; a.b2p = a.W2p + 2 * (size_t)g.FCpad * g.CP;
    a.fm = fm;
    dynca_tiles(g, a.tiles_x, a.tiles_y, a.n_tiles);
    const size_t smem = dynca_fwd_smem_floats(g) * sizeof(float);
    int occ = smem <= 113 * 1024 ? 2 : 1;
    int grid = a.n_tiles < nca_num_sms() * occ ? a.n_tiles : nca_num_sms() * occ;
    if (g.ns == 2) {
        int rc = dynca_set_smem(dynca_fwd_f32_kernel<2>, smem); if (rc) return rc;
        dynca_fwd_f32_kernel<2><<<grid, DT_THREADS, smem, s>>>(a);
    } else {
        int rc = dynca_set_smem(dynca_fwd_f32_kernel<1>, smem); if (rc) return rc;
        dynca_fwd_f32_kernel<1><<<grid, DT_THREADS, smem, s>>>(a);
    }
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int dynca_f32_perceive(const DyncaGeom& g, const float* x, const float* cond, float* z, cudaStream_t s) {
    DyncaPerceiveArgs a;
    a.g = g; a.x = x; a.cond = cond; a.z = z;
    dynca_tiles(g, a.tiles_x, a.tiles_y, a.n_tiles);
    const size_t smem = ((size_t)dynca_stage_floats(g) + (size_t)g.Ppad * DT_TMS) * sizeof(float);
    int grid = a.n_tiles < nca_num_sms() * 2 ? a.n_tiles : nca_num_sms() * 2;
    if (g.ns == 2) {
        int rc = dynca_set_smem(dynca_perceive_kernel<2>, smem); if (rc) return rc;
        dynca_perceive_kernel<2><<<grid, DT_THREADS, smem, s>>>(a);
    } else {
        int rc = dynca_set_smem(dynca_perceive_kernel<1>, smem); if (rc) return rc;
        dynca_perceive_kernel<1><<<grid, DT_THREADS, smem, s>>>(a);
    }
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int dynca_f32_backward_step(const DyncaGeom& g, const float* wsW, float* wsG, const float* x_in, const float* g_next,
                            const float* g_tap, int tap_c, float tap_scale, float* g_out, const float* cond,
                            const FireMask& fm, cudaStream_t s) {
    DyncaBwdArgs a;
    a.g = g; a.x_in = x_in; a.g_next = g_next; a.g_tap = g_tap; a.tap_c = tap_c; a.tap_scale = tap_scale;
    a.g_out = g_out; a.cond = cond;
    a.W1p = wsW; a.W2t = a.W1p + (size_t)g.Ppad * g.FCpad + (size_t)g.FCpad * g.CP;
    a.gW1p = wsG; a.gW2p = a.gW1p + (size_t)g.Ppad * g.FCpad; a.gb2p = a.gW2p + (size_t)g.FCpad * g.CP;
    a.fm = fm;
    dynca_tiles(g, a.tiles_x, a.tiles_y, a.n_tiles);
    const size_t smem = dynca_bwd_smem_floats(g) * sizeof(float);
    int grid = a.n_tiles < nca_num_sms() ? a.n_tiles : nca_num_sms();
    if (g.ns == 2) {
        int rc = dynca_set_smem(dynca_bwd_f32_kernel<2>, smem); if (rc) return rc;
        dynca_bwd_f32_kernel<2><<<grid, DT_THREADS, smem, s>>>(a);
    } else {
        int rc = dynca_set_smem(dynca_bwd_f32_kernel<1>, smem); if (rc) return rc;
        dynca_bwd_f32_kernel<1><<<grid, DT_THREADS, smem, s>>>(a);
    }
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int dynca_f32_unpack_grads(const DyncaGeom& g, const float* wsG, const NcaDyncaWeightGrads* gw, cudaStream_t s) {
    const float* gW1p = wsG; const float* gW2p = gW1p + (size_t)g.Ppad * g.FCpad; const float* gb2p = gW2p + (size_t)g.FCpad * g.CP;
    dynca_unpack_grads_kernel<<<32, 256, 0, s>>>(g, gW1p, gW2p, gb2p, gw->w1, gw->b1, gw->w2, gw->b2);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int nca_edge_extract_launch(int B, int H, int W, const float* img, int tanh_transform, float* out, cudaStream_t s) {
    size_t n = (size_t)B * H * W;
    int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    nca_edge_extract_kernel<<<grid, 256, 0, s>>>(B, H, W, img, tanh_transform, out);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int nca_philox_mask_launch(int B, int H, int W, float rate, int enc, uint64_t seed, int t0, const uint32_t* t0_dev, int T, float* out, cudaStream_t s) {
    size_t n = (size_t)T * B * H * W;
    int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    nca_philox_mask_kernel<<<grid, 256, 0, s>>>(B, H, W, (unsigned long long)nca_fire_threshold(rate, enc), enc,
                                                (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32), t0, t0_dev, T, out);
    NCA_LAUNCH_OK();
    return NCA_OK;
}
