// Tile geometry + the perception stage shared by the DyNCA kernels (fp32 CUDA-core path, perceive-only
// kernel, and the BPTT kernel).  One CTA owns a TH x TW = 4 x 32 block of cells of one sample: a warp is one
// image row of 32 consecutive cells, so global rows are 128-byte coalesced and every shared-memory stencil
// read is conflict free.
#pragma once
#include "dynca_common.cuh"

#define DT_TH 4
#define DT_TW 32
#define DT_TM 128            // cells per tile
#define DT_TMS 132           // smem row stride of [k][cell] matrices (== 4 mod 32: conflict-free float4 rows)
#define DT_XS (DT_TW + 2)    // fine state tile width incl. 1-cell ring
#define DT_XR (DT_TH + 2)
#define DT_PCH (DT_TH / 2 + 2)   // coarse perception rows feeding the bilinear x2 of one tile
#define DT_PCW (DT_TW / 2 + 2)
#define DT_CXH (DT_PCH + 2)      // coarse state rows (perception rows + ring)
#define DT_CXW (DT_PCW + 2)
#define DT_THREADS 256

// smem (floats) the perception stage needs besides sZ
__host__ __device__ static inline int dynca_stage_floats(const DyncaGeom& g) {
    int n = g.C * DT_XR * DT_XS;
    if (g.ns == 2) n += g.C * DT_CXH * DT_CXW + 4 * g.C * DT_PCH * DT_PCW;
    return (n + 3) / 4 * 4;
}

struct DyncaTile {
    int b, y0, x0;
};
__device__ __forceinline__ DyncaTile dynca_tile_of(int t, int tiles_x, int tiles_y) {
    DyncaTile r;
    r.x0 = (t % tiles_x) * DT_TW;
    t /= tiles_x;
    r.y0 = (t % tiles_y) * DT_TH;
    r.b = t / tiles_y;
    return r;
}

// 3x3 cross-correlation filters of ExtraChannels/models/dynca.py:63-69 applied to a 3x3 neighbourhood v
__device__ __forceinline__ void dynca_filters(const float v[3][3], float& sx, float& sy, float& lap) {
    sx = (v[0][2] - v[0][0]) + 2.0f * (v[1][2] - v[1][0]) + (v[2][2] - v[2][0]);
    sy = (v[2][0] - v[0][0]) + 2.0f * (v[2][1] - v[0][1]) + (v[2][2] - v[0][2]);
    lap = (v[0][0] + v[0][2] + v[2][0] + v[2][2]) + 2.0f * (v[0][1] + v[1][0] + v[1][2] + v[2][1]) - 12.0f * v[1][1];
}
// tap weights as tables (used by the transposed stencil)
__device__ __forceinline__ float dynca_tap_sx(int a, int b) { return (b == 1) ? 0.0f : ((b == 2 ? 1.0f : -1.0f) * (a == 1 ? 2.0f : 1.0f)); }
__device__ __forceinline__ float dynca_tap_sy(int a, int b) { return (a == 1) ? 0.0f : ((a == 2 ? 1.0f : -1.0f) * (b == 1 ? 2.0f : 1.0f)); }
__device__ __forceinline__ float dynca_tap_lap(int a, int b) { return (a == 1 && b == 1) ? -12.0f : ((a == 1 || b == 1) ? 2.0f : 1.0f); }

// Stage the state of sample t.b around the tile into shared memory:
//   sX  [C][DT_XR][DT_XS]    fine cells incl. the 1-cell ring, padding mode applied
//   sXc [C][DT_CXH][DT_CXW]  (NS == 2) 2x2-mean coarse cells incl. ring, padding applied on the coarse grid
//   sCP [4C][DT_PCH][DT_PCW] (NS == 2) coarse perception [id|sx|sy|lap] feeding the bilinear x2 of the tile
// Ends with __syncthreads().
template <int NS, int NT>
__device__ __forceinline__ void dynca_stage_tile(const DyncaGeom& g, const float* __restrict__ x, const DyncaTile& t,
                                                 float* __restrict__ sStage) {
    const int tid = threadIdx.x;
    const int C = g.C, H = g.H, W = g.W;
    const size_t plane = (size_t)H * W;
    const float* xb = x + (size_t)t.b * C * plane;
    float* sX = sStage;
    float* sXc = sX + C * DT_XR * DT_XS;
    float* sCP = sXc + C * DT_CXH * DT_CXW;

    for (int i = tid; i < C * DT_XR * DT_XS; i += NT) {
        int q = i % DT_XS, r = (i / DT_XS) % DT_XR, c = i / (DT_XS * DT_XR);
        int iy = nca_padmap(t.y0 - 1 + r, H, g.pad), ix = nca_padmap(t.x0 - 1 + q, W, g.pad);
        sX[i] = (iy >= 0 && ix >= 0) ? __ldg(xb + c * plane + (size_t)iy * W + ix) : 0.0f;
    }
    if (NS == 2) {
        const int Hc = H >> 1, Wc = W >> 1;
        const int cyp = (t.y0 >> 1) - 2, cxp = (t.x0 >> 1) - 2;   // coarse padded coordinate of sXc[.][0][0]
        for (int i = tid; i < C * DT_CXH * DT_CXW; i += NT) {
            int q = i % DT_CXW, r = (i / DT_CXW) % DT_CXH, c = i / (DT_CXW * DT_CXH);
            int qy = nca_padmap(cyp + r, Hc, g.pad), qx = nca_padmap(cxp + q, Wc, g.pad);
            float v = 0.0f;
            if (qy >= 0 && qx >= 0) {
                const float* p = xb + c * plane + (size_t)(2 * qy) * W + 2 * qx;
                v = 0.25f * (((__ldg(p) + __ldg(p + 1)) + __ldg(p + W)) + __ldg(p + W + 1));   // dynca.py:73-77
            }
            sXc[i] = v;
        }
    }
    __syncthreads();
    if (NS == 2) {
        for (int i = tid; i < C * DT_PCH * DT_PCW; i += NT) {
            int q = i % DT_PCW, r = (i / DT_PCW) % DT_PCH, c = i / (DT_PCW * DT_PCH);
            float v[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int bb = 0; bb < 3; ++bb) v[a][bb] = sXc[(c * DT_CXH + r + a) * DT_CXW + q + bb];
            float sx, sy, lap;
            dynca_filters(v, sx, sy, lap);
            const int o = r * DT_PCW + q, ps = DT_PCH * DT_PCW;
            sCP[(0 * C + c) * ps + o] = v[1][1];
            sCP[(1 * C + c) * ps + o] = sx;
            sCP[(2 * C + c) * ps + o] = sy;
            sCP[(3 * C + c) * ps + o] = lap;
        }
        __syncthreads();
    }
}

// bilinear x2 taps of the in-image cell (gy, gx) in sCP-local coordinates
struct DyncaUp {
    int by, bx;
    float wy0, wy1, wx0, wx1;
};
__device__ __forceinline__ DyncaUp dynca_up_of(const DyncaGeom& g, const DyncaTile& t, int gy, int gx) {
    DyncaUp u;
    dynca_up_taps(gy, g.H >> 1, u.by, u.wy0, u.wy1);
    dynca_up_taps(gx, g.W >> 1, u.bx, u.wx0, u.wx1);
    u.by -= (t.y0 >> 1) - 1;
    u.bx -= (t.x0 >> 1) - 1;
    return u;
}
// the 4 perception values [id, sx, sy, lap] of channel c at tile cell (py, px) (in-image), averaged over scales
template <int NS>
__device__ __forceinline__ void dynca_cell_percept(const DyncaGeom& g, const float* __restrict__ sStage, const DyncaUp& u,
                                                   int c, int py, int px, float f[4]) {
    const int C = g.C;
    const float* sX = sStage;
    const float* sCP = sStage + C * DT_XR * DT_XS + C * DT_CXH * DT_CXW;
    float v[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) v[a][bb] = sX[(c * DT_XR + py + a) * DT_XS + px + bb];
    f[0] = v[1][1];
    dynca_filters(v, f[1], f[2], f[3]);
    if (NS == 2) {
        const int ps = DT_PCH * DT_PCW;
        const float* cp = sCP + c * ps + u.by * DT_PCW + u.bx;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float* q = cp + k * C * ps;
            float up = u.wy0 * (u.wx0 * q[0] + u.wx1 * q[1]) + u.wy1 * (u.wx0 * q[DT_PCW] + u.wx1 * q[DT_PCW + 1]);
            f[k] = (f[k] + up) * g.s0;
        }
    }
}

// Stage x and build the perception matrix
//   sZ[k][m], k < Ppad, m = py*32+px:  rows [id C | sobel_x C | sobel_y C | lap C | cond cc | 1 | 0..]
// (DyNCA.perceive_multiscale, dynca.py:98-111).  Cells outside the image get z = 0.
// sStage: dynca_stage_floats(g) floats of scratch.  Ends with __syncthreads().
template <int NS>
__device__ __forceinline__ void dynca_perceive_tile(const DyncaGeom& g, const float* __restrict__ x,
                                                    const float* __restrict__ cond, const DyncaTile& t,
                                                    float* __restrict__ sStage, float* __restrict__ sZ) {
    const int tid = threadIdx.x;
    const int C = g.C, H = g.H, W = g.W;
    dynca_stage_tile<NS, DT_THREADS>(g, x, t, sStage);
    {
        const int m = tid & (DT_TM - 1), half = tid >> 7;
        const int py = m >> 5, px = m & 31;
        const int gy = t.y0 + py, gx = t.x0 + px;
        const bool inimg = gy < H && gx < W;
        DyncaUp u = {};
        if (NS == 2 && inimg) u = dynca_up_of(g, t, gy, gx);
        for (int c = half; c < C; c += 2) {
            float f[4] = {0.f, 0.f, 0.f, 0.f};
            if (inimg) dynca_cell_percept<NS>(g, sStage, u, c, py, px, f);
            const float f0 = f[0], f1 = f[1], f2 = f[2], f3 = f[3];
            sZ[(0 * C + c) * DT_TMS + m] = f0;
            sZ[(1 * C + c) * DT_TMS + m] = f1;
            sZ[(2 * C + c) * DT_TMS + m] = f2;
            sZ[(3 * C + c) * DT_TMS + m] = f3;
        }
        // cond rows, the constant-1 row and the zero padding rows
        for (int k = 4 * C + half; k < g.Ppad; k += 2) {
            float v = 0.0f;
            if (inimg) {
                int i = k - 4 * C;
                if (i < g.cc) {
                    if (g.cond_kind == NCA_COND_CPE) v = (i == 0) ? dynca_cpe(gy, H, g.cpe_oh) : dynca_cpe(gx, W, g.cpe_ow);
                    else v = __ldg(cond + ((size_t)(t.b * g.cc + i) * H + gy) * W + gx);
                } else if (i == g.cc) v = 1.0f;
            }
            sZ[k * DT_TMS + m] = v;
        }
    }
    __syncthreads();
}

// gradient arriving at states[t+1]: what later steps propagated plus an optional rgb tap (tap_scale * g_tap on the
// first tap_c channels)
__device__ __forceinline__ float dynca_gnext(const float* __restrict__ g_next, const float* __restrict__ g_tap, int tap_c,
                                             float tap_scale, int C, int b, int c, size_t pix, size_t plane) {
    float v = g_next ? __ldg(g_next + ((size_t)b * C + c) * plane + pix) : 0.0f;
    if (g_tap && c < tap_c) v = fmaf(tap_scale, __ldg(g_tap + ((size_t)b * tap_c + c) * plane + pix), v);
    return v;
}

// row of the [k][cell] perception-gradient matrix that holds (channel c, filter f):
//   KPERM == false: reference order  f*C + c            (fp32 path)
//   KPERM == true : tensor-core order 8*(c/2) + 4*(c%2) + f   (bf16 path, see dynca_bf16.cu)
template <bool KPERM>
__device__ __forceinline__ int dynca_krow(int C, int c, int f) {
    return KPERM ? (8 * (c >> 1) + 4 * (c & 1) + f) : (f * C + c);
}

// Transposed perception of one tile.  sZ holds s0 * g_z as [k][cell] (row stride DT_TMS); the result
//   g_t = g_{t+1} (+tap) + Perceive^T(g_z)
// is accumulated with red.add into g_out (zeroed by the caller): contributions to the tile's ring land in the
// neighbouring tiles' cells (or wrap / fold back according to the padding mode).  sG: 4C*DT_PCH*DT_PCW floats of
// scratch (NS == 2).  Caller must __syncthreads() before (sZ complete) and after (smem reuse).
//
// Work items are (output position, channel parity): the validity mask and source offset of the 9 taps depend on the
// position only, so they are computed once and the channel loop is branch-free (3 LDS + 3 FFMA per non-zero tap).
template <bool KPERM>
__device__ __forceinline__ float dynca_stencil_t(const float* __restrict__ S, int C, int c, int plane_stride, bool kperm_rows,
                                                 const int off[9], const float mk[9]) {
    // rows of the three filter planes of channel c
    const float* zx = S + (kperm_rows ? dynca_krow<KPERM>(C, c, 1) : (C + c)) * plane_stride;
    const float* zy = S + (kperm_rows ? dynca_krow<KPERM>(C, c, 2) : (2 * C + c)) * plane_stride;
    const float* zl = S + (kperm_rows ? dynca_krow<KPERM>(C, c, 3) : (3 * C + c)) * plane_stride;
    float v = 0.0f;
#pragma unroll
    for (int aa = 0; aa < 3; ++aa)
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) {
            const int k = aa * 3 + bb;
            float tv = dynca_tap_lap(aa, bb) * zl[off[k]];
            if (bb != 1) tv = fmaf(dynca_tap_sx(aa, bb), zx[off[k]], tv);
            if (aa != 1) tv = fmaf(dynca_tap_sy(aa, bb), zy[off[k]], tv);
            v = fmaf(mk[k], tv, v);
        }
    return v;
}

template <int NS, int NT, bool KPERM>
__device__ __forceinline__ void dynca_scatter_tile(const DyncaGeom& g, const DyncaTile& t, const float* __restrict__ sZ,
                                                   float* __restrict__ sU, float* __restrict__ g_out,
                                                   const float* __restrict__ g_next, const float* __restrict__ g_tap,
                                                   int tap_c, float tap_scale) {
    const int tid = threadIdx.x;
    const int C = g.C, H = g.H, W = g.W;
    const size_t plane = (size_t)H * W;
    float* gob = g_out + (size_t)t.b * C * plane;
    // ---- scale 0: transposed 3x3 stencils over the tile + ring, residual pass-through, red.add ----
    for (int item = tid; item < 2 * DT_XR * DT_XS; item += NT) {
        const int hh = item >= DT_XR * DT_XS ? 1 : 0;
        const int pos = item - hh * DT_XR * DT_XS;
        const int rx = pos % DT_XS, ry = pos / DT_XS;
        const int yy = t.y0 - 1 + ry, xx = t.x0 - 1 + rx;     // padded coordinate of this position
        if (yy > H || xx > W) continue;
        const int iy = nca_padmap(yy, H, g.pad), ix = nca_padmap(xx, W, g.pad);
        if (iy < 0 || ix < 0) continue;
        int off[9];
        float mk[9];
#pragma unroll
        for (int aa = 0; aa < 3; ++aa)
#pragma unroll
            for (int bb = 0; bb < 3; ++bb) {
                const int py = ry - aa, px = rx - bb;
                const bool ok = py >= 0 && py < DT_TH && px >= 0 && px < DT_TW;
                off[aa * 3 + bb] = ok ? py * DT_TW + px : 0;
                mk[aa * 3 + bb] = ok ? 1.0f : 0.0f;
            }
        // in-tile AND in-image cells also get the identity tap and the residual pass-through; in-tile cells
        // beyond the image edge are ordinary pad positions of the last image row / column
        const bool interior = ry >= 1 && ry <= DT_TH && rx >= 1 && rx <= DT_TW && yy < H && xx < W;
        const size_t pix = (size_t)iy * W + ix;
        for (int c = hh; c < C; c += 2) {
            float v = dynca_stencil_t<KPERM>(sZ, C, c, DT_TMS, true, off, mk);
            if (interior) {
                v += sZ[dynca_krow<KPERM>(C, c, 0) * DT_TMS + (ry - 1) * DT_TW + (rx - 1)];
                v += dynca_gnext(g_next, g_tap, tap_c, tap_scale, C, t.b, c, pix, plane);
            }
            atomicAdd(gob + c * plane + pix, v);
        }
    }
    if (NS == 2) {
        // ---- scale 1: Up^T (bilinear x2) -> coarse g_percept, stencil^T on the coarse grid, Down^T ----
        const int Hc = H >> 1, Wc = W >> 1;
        const int cy0 = (t.y0 >> 1) - 1, cx0 = (t.x0 >> 1) - 1;
        constexpr int ps = DT_PCH * DT_PCW;
        constexpr int NG = NT / ps;          // plane groups working in parallel
        float* sG = sU;                      // [4C][PCH*PCW], reference plane order f*C + c
        if (tid < NG * ps) {
            const int cell = tid % ps, grp = tid / ps;
            const int pr = cell / DT_PCW, pq = cell % DT_PCW;
            const int qy = cy0 + pr, qx = cx0 + pq;
            const bool cell_ok = qy >= 0 && qy < Hc && qx >= 0 && qx < Wc;
            float wy[4], wx[4];
            int oy[4], ox[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int fy = 2 * qy - 1 + i, py = fy - t.y0;
                const bool oky = cell_ok && py >= 0 && py < DT_TH && fy < H;
                wy[i] = oky ? dynca_up_weight(fy, qy, Hc) : 0.0f;
                oy[i] = oky ? py * DT_TW : 0;
                const int fx = 2 * qx - 1 + i, px = fx - t.x0;
                const bool okx = cell_ok && px >= 0 && px < DT_TW && fx < W;
                wx[i] = okx ? dynca_up_weight(fx, qx, Wc) : 0.0f;
                ox[i] = okx ? px : 0;
            }
            for (int k = grp; k < 4 * C; k += NG) {
                const float* zr = sZ + dynca_krow<KPERM>(C, k % C, k / C) * DT_TMS;
                float v = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float* r = zr + oy[i];
                    v = fmaf(wy[i], fmaf(wx[0], r[ox[0]], fmaf(wx[1], r[ox[1]], fmaf(wx[2], r[ox[2]], wx[3] * r[ox[3]]))), v);
                }
                sG[k * ps + cell] = v;
            }
        }
        __syncthreads();
        for (int item = tid; item < 2 * DT_CXH * DT_CXW; item += NT) {
            const int hh = item >= DT_CXH * DT_CXW ? 1 : 0;
            const int pos = item - hh * DT_CXH * DT_CXW;
            const int rx = pos % DT_CXW, ry = pos / DT_CXW;
            const int yy = cy0 - 1 + ry, xx = cx0 - 1 + rx;   // coarse padded coordinate
            if (yy < -1 || xx < -1 || yy > Hc || xx > Wc) continue;
            const int qy = nca_padmap(yy, Hc, g.pad), qx = nca_padmap(xx, Wc, g.pad);
            if (qy < 0 || qx < 0) continue;
            int off[9];
            float mk[9];
#pragma unroll
            for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                for (int bb = 0; bb < 3; ++bb) {
                    const int pr = ry - aa, pq = rx - bb;
                    const bool ok = pr >= 0 && pr < DT_PCH && pq >= 0 && pq < DT_PCW;
                    off[aa * 3 + bb] = ok ? pr * DT_PCW + pq : 0;
                    mk[aa * 3 + bb] = ok ? 1.0f : 0.0f;
                }
            const bool ident = ry >= 1 && ry <= DT_PCH && rx >= 1 && rx <= DT_PCW;
            float* p0 = gob + (size_t)(2 * qy) * W + 2 * qx;
            for (int c = hh; c < C; c += 2) {
                float v = dynca_stencil_t<KPERM>(sG, C, c, ps, false, off, mk);
                if (ident) v += sG[c * ps + (ry - 1) * DT_PCW + (rx - 1)];
                v *= 0.25f;
                float* p = p0 + c * plane;
                atomicAdd(p, v); atomicAdd(p + 1, v); atomicAdd(p + W, v); atomicAdd(p + W + 1, v);
            }
        }
    }
}

// GEMM1: acc[i][jj] = sum_k sZ[k][pix(i)] * sW1[k][ty*8+jj], k <= P (bias row included).
// pix(i) = tx*4+i (i<4) | 64+tx*4+(i-4);  tx = tid & 15, ty = tid >> 4.  Threads with ty*8 >= FCpad idle.
__device__ __forceinline__ void dynca_gemm1(const DyncaGeom& g, const float* __restrict__ sZ,
                                            const float* __restrict__ sW1, float acc[8][8]) {
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    if (ty * 8 >= g.FCpad) return;
    const float* zp = sZ + tx * 4;
    const float* wp = sW1 + ty * 8;
    const int kEnd = g.P + 1;
#pragma unroll 2
    for (int k = 0; k < kEnd; ++k) {
        float4 za = *reinterpret_cast<const float4*>(zp + k * DT_TMS);
        float4 zb = *reinterpret_cast<const float4*>(zp + k * DT_TMS + 64);
        float4 wa = *reinterpret_cast<const float4*>(wp + k * g.FCpad);
        float4 wb = *reinterpret_cast<const float4*>(wp + k * g.FCpad + 4);
        const float z[8] = {za.x, za.y, za.z, za.w, zb.x, zb.y, zb.z, zb.w};
        const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(z[i], w[j], acc[i][j]);
    }
}
