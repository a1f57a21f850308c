// Register-blocked transposed perception for the tensor-core BPTT kernel.
//
// Input: s0 * g_z of one 4x32 tile as zero-padded planes in shared memory,
//   sP[row(c,f)][py][SP2_S], the value of tile cell (py, px) at column px + 1; column 0 and columns 33..39 are zero,
// so a thread that produces 4 adjacent outputs (global x aligned to 4 -> one red.global.add.v4.f32) reads, per source
// row and plane, two aligned float4 (columns 4b .. 4b+7) and needs no bounds logic in x.  Rows are handled by
// giving every warp a single output row ry, so "is source row ry-a inside the tile" is warp-uniform.
// Coarse scale: Up^T gathers into zero-padded coarse planes sG[f*C+c][pr][SG2_S] (18 data columns at 0..17), the same
// blocked stencil^T runs on the coarse grid and Down^T replicates each coarse value to its 2x2 fine cells
// (0.25 each) with four v4 reductions per thread.
#pragma once
#include "dynca_tile.cuh"

#define SP2_S 40
#define SP2_PLANE (DT_TH * SP2_S)        // 160 floats
#define SG2_S 24
#define SG2_PLANE (DT_PCH * SG2_S)       // 96 floats

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// out[i] += sum over the 3 source columns (bb) of one source row (tap row aa) for 4 adjacent outputs.
// v[0..7]: plane values at padded columns base .. base+7; output i, tap bb reads v[i + 2 - bb]
template <int AA>
__device__ __forceinline__ void sp2_row_taps(const float* __restrict__ px_, const float* __restrict__ py_, const float* __restrict__ pl_,
                                             float out[4]) {
    const float4 xa = *reinterpret_cast<const float4*>(px_), xb = *reinterpret_cast<const float4*>(px_ + 4);
    const float4 ya = *reinterpret_cast<const float4*>(py_), yb = *reinterpret_cast<const float4*>(py_ + 4);
    const float4 la = *reinterpret_cast<const float4*>(pl_), lb = *reinterpret_cast<const float4*>(pl_ + 4);
    const float vx[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
    const float vy[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
    const float vl[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) {
            const int j = i + 2 - bb;
            float tv = dynca_tap_lap(AA, bb) * vl[j];
            if (bb != 1) tv = fmaf(dynca_tap_sx(AA, bb), vx[j], tv);
            if (AA != 1) tv = fmaf(dynca_tap_sy(AA, bb), vy[j], tv);
            out[i] += tv;
        }
}

// generic single output (ring columns, ragged tiles): out = sum_{aa,bb valid} taps * plane[(ry-aa)][px = rx - bb]
// planes addressed as P[f][py * S + px + off]; px valid in [0, ncols)
__device__ __forceinline__ float sp2_point(const float* __restrict__ px_, const float* __restrict__ py_, const float* __restrict__ pl_,
                                           int S, int off, int nrows, int ncols, int ry, int rx) {
    float v = 0.0f;
#pragma unroll
    for (int aa = 0; aa < 3; ++aa) {
        const int py = ry - aa;
        if (py < 0 || py >= nrows) continue;
#pragma unroll
        for (int bb = 0; bb < 3; ++bb) {
            const int px = rx - bb;
            if (px < 0 || px >= ncols) continue;
            const int o = py * S + px + off;
            float tv = dynca_tap_lap(aa, bb) * pl_[o];
            if (bb != 1) tv = fmaf(dynca_tap_sx(aa, bb), px_[o], tv);
            if (aa != 1) tv = fmaf(dynca_tap_sy(aa, bb), py_[o], tv);
            v += tv;
        }
    }
    return v;
}

// sP: padded fine planes, rows indexed by dynca_krow<true>; sG: scratch for 4C coarse planes (SG2_PLANE floats each);
// sW: 88 floats of scratch for the bilinear weight tables.
// Caller: __syncthreads() before (sP complete incl. zero columns) and after.
template <int NS, int NT>
__device__ __forceinline__ void dynca_scatter_tile_v2(const DyncaGeom& g, const DyncaTile& t, const float* __restrict__ sP,
                                                      float* __restrict__ sG, float* __restrict__ sW, float* __restrict__ g_out,
                                                      const float* __restrict__ g_next, const float* __restrict__ g_tap,
                                                      int tap_c, float tap_scale) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = NT / 32;
    const int C = g.C, H = g.H, W = g.W;
    const size_t plane = (size_t)H * W;
    float* gob = g_out + (size_t)t.b * C * plane;
    const bool fast = ((W & 3) == 0) && (t.x0 + DT_TW <= W);
    // ---------------- scale 0 ----------------
    if (fast) {
        // blocked interior columns: warp-item = (ry, group of 4 channels); lane = (channel in group, block b of 4 columns)
        const int wpr = (C + 3) >> 2;
        for (int wi = warp; wi < DT_XR * wpr; wi += NW) {
            const int ry = wi / wpr, c = (wi % wpr) * 4 + (lane >> 3), b = lane & 7;
            const int yy = t.y0 - 1 + ry;
            if (yy > H) continue;                                     // warp-uniform
            const int iy = nca_padmap(yy, H, g.pad);
            if (iy < 0 || c >= C) continue;
            const float* pxp = sP + dynca_krow<true>(C, c, 1) * SP2_PLANE + 4 * b;
            const float* pyp = sP + dynca_krow<true>(C, c, 2) * SP2_PLANE + 4 * b;
            const float* plp = sP + dynca_krow<true>(C, c, 3) * SP2_PLANE + 4 * b;
            float out[4] = {0.f, 0.f, 0.f, 0.f};
            if (ry >= 0 && ry < DT_TH) sp2_row_taps<0>(pxp + ry * SP2_S, pyp + ry * SP2_S, plp + ry * SP2_S, out);
            if (ry >= 1 && ry < DT_TH + 1) sp2_row_taps<1>(pxp + (ry - 1) * SP2_S, pyp + (ry - 1) * SP2_S, plp + (ry - 1) * SP2_S, out);
            if (ry >= 2 && ry < DT_TH + 2) sp2_row_taps<2>(pxp + (ry - 2) * SP2_S, pyp + (ry - 2) * SP2_S, plp + (ry - 2) * SP2_S, out);
            const size_t pix = (size_t)iy * W + t.x0 + 4 * b;
            if (ry >= 1 && ry <= DT_TH && yy < H) {                    // identity tap + residual pass-through
                const float* pid = sP + dynca_krow<true>(C, c, 0) * SP2_PLANE + (ry - 1) * SP2_S + 4 * b + 1;
                float4 gn = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g_next) gn = __ldg(reinterpret_cast<const float4*>(g_next + ((size_t)t.b * C + c) * plane + pix));
                if (g_tap && c < tap_c) {
                    const float4 tp = __ldg(reinterpret_cast<const float4*>(g_tap + ((size_t)t.b * tap_c + c) * plane + pix));
                    gn.x = fmaf(tap_scale, tp.x, gn.x); gn.y = fmaf(tap_scale, tp.y, gn.y);
                    gn.z = fmaf(tap_scale, tp.z, gn.z); gn.w = fmaf(tap_scale, tp.w, gn.w);
                }
                out[0] += pid[0] + gn.x; out[1] += pid[1] + gn.y; out[2] += pid[2] + gn.z; out[3] += pid[3] + gn.w;
            }
            red_add_v4(gob + c * plane + pix, out[0], out[1], out[2], out[3]);
        }
        // the two ring columns
        for (int it = tid; it < C * DT_XR * 2; it += NT) {
            const int side = it & 1, ry = (it >> 1) % DT_XR, c = it / (2 * DT_XR);
            const int rx = side ? DT_XS - 1 : 0;
            const int yy = t.y0 - 1 + ry, xx = t.x0 - 1 + rx;
            if (yy > H || xx > W) continue;
            const int iy = nca_padmap(yy, H, g.pad), ix = nca_padmap(xx, W, g.pad);
            if (iy < 0 || ix < 0) continue;
            const float v = sp2_point(sP + dynca_krow<true>(C, c, 1) * SP2_PLANE, sP + dynca_krow<true>(C, c, 2) * SP2_PLANE,
                                      sP + dynca_krow<true>(C, c, 3) * SP2_PLANE, SP2_S, 1, DT_TH, DT_TW, ry, rx);
            atomicAdd(gob + c * plane + (size_t)iy * W + ix, v);
        }
    } else {
        for (int it = tid; it < C * DT_XR * DT_XS; it += NT) {
            const int rx = it % DT_XS, ry = (it / DT_XS) % DT_XR, c = it / (DT_XS * DT_XR);
            const int yy = t.y0 - 1 + ry, xx = t.x0 - 1 + rx;
            if (yy > H || xx > W) continue;
            const int iy = nca_padmap(yy, H, g.pad), ix = nca_padmap(xx, W, g.pad);
            if (iy < 0 || ix < 0) continue;
            float v = sp2_point(sP + dynca_krow<true>(C, c, 1) * SP2_PLANE, sP + dynca_krow<true>(C, c, 2) * SP2_PLANE,
                                sP + dynca_krow<true>(C, c, 3) * SP2_PLANE, SP2_S, 1, DT_TH, DT_TW, ry, rx);
            if (ry >= 1 && ry <= DT_TH && rx >= 1 && rx <= DT_TW && yy < H && xx < W) {
                v += sP[dynca_krow<true>(C, c, 0) * SP2_PLANE + (ry - 1) * SP2_S + rx];
                v += dynca_gnext(g_next, g_tap, tap_c, tap_scale, C, t.b, c, (size_t)yy * W + xx, plane);
            }
            atomicAdd(gob + c * plane + (size_t)iy * W + ix, v);
        }
    }
    if (NS == 2) {
        // ---------------- scale 1 ----------------
        const int Hc = H >> 1, Wc = W >> 1;
        const int cy0 = (t.y0 >> 1) - 1, cx0 = (t.x0 >> 1) - 1;
        // Up^T, separable: T[py] = sum_i wx[pq][i] * g[py][2pq-3+i]  (x pass, zero padding supplies the out-of-tile taps),
        // sG[f*C+c][pr][pq] = sum_py wy[pr][py] * T[py] (y pass).  The bilinear weights (edge clamped, dynca.py:93-94)
        // depend on the tile only: 16 + 72 values in shared memory (sW, 88 floats).  Coarse column pq reads the 4 padded
        // columns starting at max(2pq-2, 0), i.e. tile cells px = base-1 .. base+2; dynca_up_weight is 0 off the footprint.
        float* sWy = sW;                            // [4 pr][4 py]
        float* sWx = sW + 16;                       // [18 pq][4 j]
        if (tid < 16) {
            const int fy = t.y0 + (tid & 3);
            sWy[tid] = fy < H ? dynca_up_weight(fy, cy0 + (tid >> 2), Hc) : 0.0f;
        } else if (tid >= 32 && tid < 32 + 4 * DT_PCW) {
            const int pq = (tid - 32) >> 2, j = (tid - 32) & 3;
            const int base = 2 * pq - 2 > 0 ? 2 * pq - 2 : 0;
            const int px = base - 1 + j, fx = t.x0 + px, qx = cx0 + pq;
            sWx[tid - 32] = (px >= 0 && px < DT_TW && fx < W && qx >= 0 && qx < Wc) ? dynca_up_weight(fx, qx, Wc) : 0.0f;
        }
        __syncthreads();
        for (int it = tid; it < 4 * C * DT_PCW; it += NT) {
            const int k = it / DT_PCW, pq = it % DT_PCW;
            const float* src = sP + dynca_krow<true>(C, k % C, k / C) * SP2_PLANE + (2 * pq - 2 > 0 ? 2 * pq - 2 : 0);
            const float4 wx = *reinterpret_cast<const float4*>(sWx + 4 * pq);
            float T[4];
#pragma unroll
            for (int py = 0; py < DT_TH; ++py) {
                const float2 v0 = *reinterpret_cast<const float2*>(src + py * SP2_S), v1 = *reinterpret_cast<const float2*>(src + py * SP2_S + 2);
                T[py] = fmaf(wx.x, v0.x, fmaf(wx.y, v0.y, fmaf(wx.z, v1.x, wx.w * v1.y)));
            }
            float* dst = sG + k * SG2_PLANE + pq;
#pragma unroll
            for (int pr = 0; pr < DT_PCH; ++pr) {
                const float4 wy = *reinterpret_cast<const float4*>(sWy + 4 * pr);
                dst[pr * SG2_S] = fmaf(wy.x, T[0], fmaf(wy.y, T[1], fmaf(wy.z, T[2], wy.w * T[3])));
            }
        }
        for (int it = tid; it < 4 * C * DT_PCH * (SG2_S - DT_PCW); it += NT) {   // zero columns 18..23
            const int q = it % (SG2_S - DT_PCW), r = it / (SG2_S - DT_PCW);
            sG[r * SG2_S + DT_PCW + q] = 0.0f;
        }
        __syncthreads();
        const bool cfast = ((W & 7) == 0) && (t.x0 + DT_TW <= W);
        if (cfast) {
            // blocked: coarse columns rx = 2 + 4b + i (i < 4) <-> coarse x = x0/2 + 4b + i ; planes stored at pq, px = rx - bb
            // -> padded columns [4b .. 4b+5]
            const int wpr = (C + 7) >> 3;
            for (int wi = warp; wi < DT_CXH * wpr; wi += NW) {
                const int ry = wi / wpr, c = (wi % wpr) * 8 + (lane >> 2), b = lane & 3;
                const int yy = cy0 - 1 + ry;
                if (yy < -1 || yy > Hc) continue;                      // warp-uniform
                const int qy = nca_padmap(yy, Hc, g.pad);
                if (qy < 0 || c >= C) continue;
                const float* pxp = sG + (C + c) * SG2_PLANE + 4 * b;
                const float* pyp = sG + (2 * C + c) * SG2_PLANE + 4 * b;
                const float* plp = sG + (3 * C + c) * SG2_PLANE + 4 * b;
                float out[4] = {0.f, 0.f, 0.f, 0.f};
                if (ry >= 0 && ry < DT_PCH) sp2_row_taps<0>(pxp + ry * SG2_S, pyp + ry * SG2_S, plp + ry * SG2_S, out);
                if (ry >= 1 && ry < DT_PCH + 1) sp2_row_taps<1>(pxp + (ry - 1) * SG2_S, pyp + (ry - 1) * SG2_S, plp + (ry - 1) * SG2_S, out);
                if (ry >= 2 && ry < DT_PCH + 2) sp2_row_taps<2>(pxp + (ry - 2) * SG2_S, pyp + (ry - 2) * SG2_S, plp + (ry - 2) * SG2_S, out);
                if (ry >= 1 && ry <= DT_PCH) {                         // identity: coarse cell (ry-1, rx-1), rx = 2+4b+i
                    const float* pid = sG + c * SG2_PLANE + (ry - 1) * SG2_S + 4 * b + 1;
                    out[0] += pid[0]; out[1] += pid[1]; out[2] += pid[2]; out[3] += pid[3];
                }
                float* p = gob + c * plane + (size_t)(2 * qy) * W + t.x0 + 8 * b;
                const float a0 = 0.25f * out[0], a1 = 0.25f * out[1], a2 = 0.25f * out[2], a3 = 0.25f * out[3];
                red_add_v4(p, a0, a0, a1, a1); red_add_v4(p + 4, a2, a2, a3, a3);
                red_add_v4(p + W, a0, a0, a1, a1); red_add_v4(p + W + 4, a2, a2, a3, a3);
            }
            // ring columns rx in {0, 1, 18, 19}
            for (int it = tid; it < C * DT_CXH * 4; it += NT) {
                const int k = it & 3, ry = (it >> 2) % DT_CXH, c = it / (4 * DT_CXH);
                const int rx = k < 2 ? k : DT_CXW - 4 + k;
                const int yy = cy0 - 1 + ry, xx = cx0 - 1 + rx;
                if (yy < -1 || xx < -1 || yy > Hc || xx > Wc) continue;
                const int qy = nca_padmap(yy, Hc, g.pad), qx = nca_padmap(xx, Wc, g.pad);
                if (qy < 0 || qx < 0) continue;
                float v = sp2_point(sG + (C + c) * SG2_PLANE, sG + (2 * C + c) * SG2_PLANE, sG + (3 * C + c) * SG2_PLANE,
                                    SG2_S, 0, DT_PCH, DT_PCW, ry, rx);
                if (ry >= 1 && ry <= DT_PCH && rx >= 1 && rx <= DT_PCW) v += sG[c * SG2_PLANE + (ry - 1) * SG2_S + rx - 1];
                v *= 0.25f;
                float* p = gob + c * plane + (size_t)(2 * qy) * W + 2 * qx;
                atomicAdd(p, v); atomicAdd(p + 1, v); atomicAdd(p + W, v); atomicAdd(p + W + 1, v);
            }
        } else {
            for (int it = tid; it < C * DT_CXH * DT_CXW; it += NT) {
                const int rx = it % DT_CXW, ry = (it / DT_CXW) % DT_CXH, c = it / (DT_CXW * DT_CXH);
                const int yy = cy0 - 1 + ry, xx = cx0 - 1 + rx;
                if (yy < -1 || xx < -1 || yy > Hc || xx > Wc) continue;
                const int qy = nca_padmap(yy, Hc, g.pad), qx = nca_padmap(xx, Wc, g.pad);
                if (qy < 0 || qx < 0) continue;
                float v = sp2_point(sG + (C + c) * SG2_PLANE, sG + (2 * C + c) * SG2_PLANE, sG + (3 * C + c) * SG2_PLANE,
                                    SG2_S, 0, DT_PCH, DT_PCW, ry, rx);
                if (ry >= 1 && ry <= DT_PCH && rx >= 1 && rx <= DT_PCW) v += sG[c * SG2_PLANE + (ry - 1) * SG2_S + rx - 1];
                v *= 0.25f;
                float* p = gob + c * plane + (size_t)(2 * qy) * W + 2 * qx;
                atomicAdd(p, v); atomicAdd(p + 1, v); atomicAdd(p + W, v); atomicAdd(p + W + 1, v);
            }
        }
    }
}
