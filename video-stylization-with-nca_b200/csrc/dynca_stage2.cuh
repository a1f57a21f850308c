// Asynchronous tile staging for the tensor-core kernels: the fine state tile (+1-cell ring) and, for two perception
// scales, the coarse (2x2-mean) state tile (+2-cell ring) are brought in with cp.async (16-byte copies for the
// 16-byte aligned interior, 4-byte copies for the ring columns / ragged tiles), so every load of a tile is in flight
// at once and the copy of the NEXT tile can overlap the tensor-core / epilogue / scatter phases of the current one.
// The coarse state xc [B,C,H/2,W/2] is produced once per step by dynca_coarsen_kernel (dynca.py:73-77).
#pragma once
#include "dynca_tile.cuh"

#define DS_XS 40      // fine row stride; ring-extended column q (0..33) lives at q + DS_XOFF -> interior 16 B aligned
#define DS_XOFF 3
#define DS_CS 24      // coarse row stride; column q (0..19) lives at q + DS_COFF
#define DS_COFF 2

__host__ __device__ static inline int dynca_stage2_floats(const DyncaGeom& g) {
    int n = g.C * DT_XR * DS_XS;
    if (g.ns == 2) n += g.C * DT_CXH * DS_CS + 4 * g.C * DT_PCH * DT_PCW;
    return (n + 3) / 4 * 4;
}

__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// issue the copies of tile t (no wait).  x: state [B,C,H,W]; xc: coarse state [B,C,H/2,W/2] (NS == 2)
template <int NS, int NT>
__device__ __forceinline__ void dynca_stage2_issue(const DyncaGeom& g, const float* __restrict__ x, const float* __restrict__ xc,
                                                   const DyncaTile& t, float* __restrict__ sStage) {
    const int tid = threadIdx.x;
    const int C = g.C, H = g.H, W = g.W;
    const size_t plane = (size_t)H * W;
    const float* xb = x + (size_t)t.b * C * plane;
    float* sX = sStage;
    const bool fast = ((W & 3) == 0) && (t.x0 + DT_TW <= W);
    if (fast) {
        for (int it = tid; it < C * DT_XR * 8; it += NT) {            // interior columns, 8 x 16 B per row
            const int v = it & 7, r = (it >> 3) % DT_XR, c = it / (8 * DT_XR);
            const int iy = nca_padmap(t.y0 - 1 + r, H, g.pad);
            float* dst = sX + (c * DT_XR + r) * DS_XS + DS_XOFF + 1 + 4 * v;
            if (iy >= 0) cp_async16(dst, xb + c * plane + (size_t)iy * W + t.x0 + 4 * v);
            else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int it = tid; it < C * DT_XR * 2; it += NT) {            // the two ring columns
            const int side = it & 1, r = (it >> 1) % DT_XR, c = it / (2 * DT_XR);
            const int q = side ? DT_XS - 1 : 0;
            const int iy = nca_padmap(t.y0 - 1 + r, H, g.pad), ix = nca_padmap(t.x0 - 1 + q, W, g.pad);
            float* dst = sX + (c * DT_XR + r) * DS_XS + DS_XOFF + q;
            if (iy >= 0 && ix >= 0) cp_async4(dst, xb + c * plane + (size_t)iy * W + ix);
            else *dst = 0.0f;
        }
    } else {
        for (int it = tid; it < C * DT_XR * DT_XS; it += NT) {
            const int q = it % DT_XS, r = (it / DT_XS) % DT_XR, c = it / (DT_XS * DT_XR);
            const int iy = nca_padmap(t.y0 - 1 + r, H, g.pad), ix = nca_padmap(t.x0 - 1 + q, W, g.pad);
            float* dst = sX + (c * DT_XR + r) * DS_XS + DS_XOFF + q;
            if (iy >= 0 && ix >= 0) cp_async4(dst, xb + c * plane + (size_t)iy * W + ix);
            else *dst = 0.0f;
        }
    }
    if (NS == 2) {
        const int Hc = H >> 1, Wc = W >> 1;
        const size_t cplane = (size_t)Hc * Wc;
        const float* cb = xc + (size_t)t.b * C * cplane;
        float* sXc = sX + C * DT_XR * DS_XS;
        const int cyp = (t.y0 >> 1) - 2, cxp = (t.x0 >> 1) - 2;       // coarse padded coordinate of column / row 0
        const bool cfast = ((Wc & 3) == 0) && ((t.x0 >> 1) + DT_TW / 2 <= Wc);
        if (cfast) {
            for (int it = tid; it < C * DT_CXH * 4; it += NT) {       // 16 interior coarse columns
                const int v = it & 3, r = (it >> 2) % DT_CXH, c = it / (4 * DT_CXH);
                const int qy = nca_padmap(cyp + r, Hc, g.pad);
                float* dst = sXc + (c * DT_CXH + r) * DS_CS + DS_COFF + 2 + 4 * v;
                if (qy >= 0) cp_async16(dst, cb + c * cplane + (size_t)qy * Wc + (t.x0 >> 1) + 4 * v);
                else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            for (int it = tid; it < C * DT_CXH * 4; it += NT) {       // 2 + 2 ring columns
                const int k = it & 3, r = (it >> 2) % DT_CXH, c = it / (4 * DT_CXH);
                const int q = k < 2 ? k : DT_CXW - 4 + k;
                const int qy = nca_padmap(cyp + r, Hc, g.pad), qx = nca_padmap(cxp + q, Wc, g.pad);
                float* dst = sXc + (c * DT_CXH + r) * DS_CS + DS_COFF + q;
                if (qy >= 0 && qx >= 0) cp_async4(dst, cb + c * cplane + (size_t)qy * Wc + qx);
                else *dst = 0.0f;
            }
        } else {
            for (int it = tid; it < C * DT_CXH * DT_CXW; it += NT) {
                const int q = it % DT_CXW, r = (it / DT_CXW) % DT_CXH, c = it / (DT_CXW * DT_CXH);
                const int qy = nca_padmap(cyp + r, Hc, g.pad), qx = nca_padmap(cxp + q, Wc, g.pad);
                float* dst = sXc + (c * DT_CXH + r) * DS_CS + DS_COFF + q;
                if (qy >= 0 && qx >= 0) cp_async4(dst, cb + c * cplane + (size_t)qy * Wc + qx);
                else *dst = 0.0f;
            }
        }
    }
    cp_async_commit();
}

// wait for the copies, then (NS == 2) build the coarse perception sCP [4C][DT_PCH][DT_PCW].  Ends with __syncthreads().
template <int NS, int NT>
__device__ __forceinline__ void dynca_stage2_finish(const DyncaGeom& g, float* __restrict__ sStage) {
    cp_async_wait_all();
    __syncthreads();
    if (NS == 2) {
        const int C = g.C;
        const float* sXc = sStage + C * DT_XR * DS_XS;
        float* sCP = sStage + C * DT_XR * DS_XS + C * DT_CXH * DS_CS;
        for (int i = threadIdx.x; i < C * DT_PCH * DT_PCW; i += NT) {
            const int q = i % DT_PCW, r = (i / DT_PCW) % DT_PCH, c = i / (DT_PCW * DT_PCH);
            float v[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int bb = 0; bb < 3; ++bb) v[a][bb] = sXc[(c * DT_CXH + r + a) * DS_CS + DS_COFF + q + bb];
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

// the 4 perception values [id, sx, sy, lap] of channel c at tile cell (py, px) (in-image), averaged over scales
template <int NS>
__device__ __forceinline__ void dynca_cell_percept2(const DyncaGeom& g, const float* __restrict__ sStage, const DyncaUp& u,
                                                    int c, int py, int px, float f[4]) {
    const int C = g.C;
    const float* sX = sStage + DS_XOFF;
    const float* sCP = sStage + C * DT_XR * DS_XS + C * DT_CXH * DS_CS;
    float v[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) v[a][bb] = sX[(c * DT_XR + py + a) * DS_XS + px + bb];
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

// coarse state of one step: xc[b][c][qy][qx] = 2x2 mean of x (bilinear /2 on even sizes, dynca.py:73-77)
__global__ void dynca_coarsen_kernel(int BC, int H, int W, const float* __restrict__ x, float* __restrict__ xc) {
    const int Hc = H >> 1, Wc = W >> 1;
    const size_t n = (size_t)BC * Hc * Wc;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int qx = (int)(i % Wc), qy = (int)((i / Wc) % Hc);
        const size_t bc = i / ((size_t)Wc * Hc);
        const float* p = x + bc * H * W + (size_t)(2 * qy) * W + 2 * qx;
        const float2 a = *reinterpret_cast<const float2*>(p), b = *reinterpret_cast<const float2*>(p + W);
        xc[i] = 0.25f * (((a.x + a.y) + b.x) + b.y);
    }
}
