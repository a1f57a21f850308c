// DyNCA forward step with the update MLP on the 5th-gen tensor cores (NCA_PREC_BF16).
// Reference semantics: ExtraChannels/models/dynca.py:113-123 (same step as dynca_f32.cu).
//
// One CTA = 128 threads = one 4x32 tile of cells; thread r owns cell r (TMEM lane r, row r of both A operands).
//   perception (CUDA cores, fp32)  ->  A1 [128 x K1] bf16 in shared memory, UMMA canonical K-major layout
//   GEMM1  D1[128 x fc]  = A1 . W1^T      tcgen05.mma kind::f16, M=128, N=fc, K1/16 instructions, fp32 accum in TMEM
//   epilogue 1: tcgen05.ld D1 -> relu -> bf16 -> A2 [128 x fc] in shared memory
//   GEMM2  D2[128 x 16]  = A2 . W2^T      M=128, N=16, fc/16 instructions
//   epilogue 2: tcgen05.ld D2 -> + b2 -> * fire mask -> + x -> coalesced NCHW store
// K order of A1 / W1 is permuted so that one thread produces whole 16-byte rows of core matrices:
//   k' = 8*cp + 4*h + f  <->  channel c = 2*cp + h, filter f (id, sobel_x, sobel_y, lap);
//   chunk cp = ceil(C/2): [cond_0 .. cond_{cc-1}, 1, 1, 0 ..]  (the two constant-1 columns carry b1 split into
//   bf16 hi + lo parts, so the bias keeps ~16 bits of mantissa).
// NCA_PREC_F16X3 (template parameter X3): every operand is carried as an fp16 hi + lo pair and every product A.B is issued as the
// three MMAs Ah.Bh + Al.Bh + Ah.Bl into the same fp32 accumulator - ~22 significant bits per operand, which meets the fp32-grade
// parity bars (state 1e-5, gradients 1e-4) on the tensor cores; perception, fire mask, residual and all reductions are fp32 anyway.
// Range: weight images are scaled by 2^8 and gradient operands by a per-launch power of two (dynca_tc_common.cuh); the epilogues
// undo both exactly.
// Operand smem layout (no swizzle): element (row, k) at  (k/8)*LBO + (row/8)*128 + (row%8)*16 + (k%8)*2  bytes,
// i.e. 8x8 core matrices of 128 contiguous bytes, SBO = 128 (next 8 rows), LBO = rows*16 (next 8 k).
#include <cuda_bf16.h>
#include "dynca_stage2.cuh"
#include "dynca_scatter2.cuh"
#include "nca_internal.h"

#define BT_THREADS 128

#include "dynca_tc_common.cuh"

// ---- weight packing: fp32 reference layout -> bf16 UMMA operand images -------------------------------
__global__ void dynca_bf16_prep_kernel(DyncaGeom g, Bf16Geom bg, const float* __restrict__ w1, const float* __restrict__ b1,
                                       const float* __restrict__ w2, const float* __restrict__ b2,
                                       __nv_bfloat16* __restrict__ B1, __nv_bfloat16* __restrict__ B2, float* __restrict__ b2p,
                                       __nv_bfloat16* __restrict__ B1lo, __nv_bfloat16* __restrict__ B2lo) {
    const int n1 = bg.K1 * g.fc, n2 = g.fc * 16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2 + 16; i += gridDim.x * blockDim.x) {
        if (i < n1) {
            const int kp = i / g.fc, j = i % g.fc;            // (k', hidden unit)
            const int kc = kp >> 3, s = kp & 7;
            float v = 0.0f;
            if (kc < bg.npairs) {
                const int c = 2 * kc + (s >> 2), f = s & 3;
                if (c < g.C) v = w1[j * g.P + f * g.C + c];
            } else if (kc == bg.npairs) {
                const int src = dynca_cond_slot_src(g.cc, s);
                if (src >= 0) v = w1[j * g.P + 4 * g.C + src];
                else if (src == -2) v = __bfloat162float(__float2bfloat16_rn(b1[j]));
                else if (src == -3) v = b1[j] - __bfloat162float(__float2bfloat16_rn(b1[j]));
            }
            const size_t off = (size_t)kc * (g.fc / 8) * 64 + (size_t)(j >> 3) * 64 + (j & 7) * 8 + s;   // in elements
            if (B1lo) {      // F16X3: fp16 hi / lo of the scaled weight
                const __half h = __float2half_rn(v * NCA_X3_WSCALE);
                reinterpret_cast<__half*>(B1)[off] = h;
                reinterpret_cast<__half*>(B1lo)[off] = __float2half_rn(v * NCA_X3_WSCALE - __half2float(h));
            } else {
                B1[off] = __float2bfloat16_rn(v);
            }
        } else if (i < n1 + n2) {
            const int e = i - n1, j = e / 16, c = e % 16;     // (hidden unit = k, channel = n)
            const float v = c < g.C ? w2[c * g.fc + j] : 0.0f;
            const size_t off = (size_t)(j >> 3) * 128 + (size_t)(c >> 3) * 64 + (c & 7) * 8 + (j & 7);
            if (B2lo) {
                const __half h = __float2half_rn(v * NCA_X3_WSCALE);
                reinterpret_cast<__half*>(B2)[off] = h;
                reinterpret_cast<__half*>(B2lo)[off] = __float2half_rn(v * NCA_X3_WSCALE - __half2float(h));
            } else {
                B2[off] = __float2bfloat16_rn(v);
            }
        } else {
            const int c = i - n1 - n2;
            b2p[c] = c < g.C ? b2[c] : 0.0f;
        }
    }
}


__global__ void dynca_bf16_prep_b1_kernel(DyncaGeom g, Bf16Geom bg, const float* __restrict__ w1, const float* __restrict__ b1,
                                          __nv_bfloat16* __restrict__ B1, __nv_bfloat16* __restrict__ B1lo) {
    const int n1 = bg.K1 * g.fc;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += gridDim.x * blockDim.x) {
        const int kp = i / g.fc, j = i % g.fc;
        const int kc = kp >> 3, s = kp & 7;
        float v = 0.0f;
        if (kc < bg.npairs) {
            const int c = 2 * kc + (s >> 2), f = s & 3;
            if (c < g.C) v = w1[j * g.P + f * g.C + c];
        } else if (kc == bg.npairs) {
            const int src = dynca_cond_slot_src(g.cc, s);
            if (src >= 0) v = w1[j * g.P + 4 * g.C + src];
            else if (src == -2) v = __bfloat162float(__float2bfloat16_rn(b1[j]));
            else if (src == -3) v = b1[j] - __bfloat162float(__float2bfloat16_rn(b1[j]));
        }
        const size_t off = (size_t)kc * (g.fc / 8) * 64 + (size_t)(j >> 3) * 64 + (j & 7) * 8 + s;
        if (B1lo) {
            const __half h = __float2half_rn(v * NCA_X3_WSCALE);
            reinterpret_cast<__half*>(B1)[off] = h;
            reinterpret_cast<__half*>(B1lo)[off] = __float2half_rn(v * NCA_X3_WSCALE - __half2float(h));
        } else {
            B1[off] = __float2bfloat16_rn(v);
        }
    }
}

// ---- forward step ------------------------------------------------------------------------------------
struct DyncaBf16Args {
    DyncaGeom g;
    Bf16Geom bg;
    const float* x_in; const float* xc; float* x_out; const float* cond;
    const __nv_bfloat16* B1; const __nv_bfloat16* B2; const float* b2p;
    const __nv_bfloat16* B1lo; const __nv_bfloat16* B2lo;      // X3: lo images
    FireMask fm;
    int tiles_x, tiles_y, n_tiles;
};

static inline size_t dynca_bf16_smem_bytes(const DyncaGeom& g, const Bf16Geom& bg, bool x3) {
    const size_t m = x3 ? 2 : 1;
    size_t stage = (size_t)dynca_stage2_floats(g) * 4;
    size_t u = stage > m * bg.a2_bytes ? stage : m * bg.a2_bytes;
    return 1024 /*alignment slack*/ + 128 /*barrier, tmem ptr, b2*/ + m * (bg.a1_bytes + bg.b1_bytes + bg.b2_bytes) + u;
}

template <int NS, bool X3>
__global__ void __launch_bounds__(BT_THREADS) dynca_fwd_bf16_kernel(const DyncaBf16Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const DyncaGeom& g = a.g;
    const Bf16Geom& bg = a.bg;
    uint8_t* base = smem_raw;   // 1024-byte aligned by declaration; no integer round-trip, so accesses stay LDS/STS
    uint64_t* bar = reinterpret_cast<uint64_t*>(base);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + 8);
    float* sB2 = reinterpret_cast<float*>(base + 64);          // 16 floats
    // X3: every operand region holds the hi image followed by the lo image
    constexpr uint32_t XM = X3 ? 2u : 1u;
    uint8_t* sA1 = base + 128;
    uint8_t* sA1l = sA1 + bg.a1_bytes;
    uint8_t* sB1 = sA1 + XM * bg.a1_bytes;
    uint8_t* sB1l = sB1 + bg.b1_bytes;
    uint8_t* sB2w = sB1 + XM * bg.b1_bytes;
    uint8_t* sB2wl = sB2w + bg.b2_bytes;
    uint8_t* sU = sB2w + XM * bg.b2_bytes;                     // stage (fp32) | A2 (bf16)
    float* sStage = reinterpret_cast<float*>(sU);
    uint8_t* sA2 = sU;
    uint8_t* sA2l = sA2 + bg.a2_bytes;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int C = g.C, H = g.H, W = g.W;
    const size_t plane = (size_t)H * W;

    // ---- one-time setup: weights -> smem, barrier, TMEM ----
    for (uint32_t i = tid; i < bg.b1_bytes / 16; i += BT_THREADS)
        reinterpret_cast<uint4*>(sB1)[i] = __ldg(reinterpret_cast<const uint4*>(a.B1) + i);
    for (uint32_t i = tid; i < bg.b2_bytes / 16; i += BT_THREADS)
        reinterpret_cast<uint4*>(sB2w)[i] = __ldg(reinterpret_cast<const uint4*>(a.B2) + i);
    if (X3) {
        for (uint32_t i = tid; i < bg.b1_bytes / 16; i += BT_THREADS)
            reinterpret_cast<uint4*>(sB1l)[i] = __ldg(reinterpret_cast<const uint4*>(a.B1lo) + i);
        for (uint32_t i = tid; i < bg.b2_bytes / 16; i += BT_THREADS)
            reinterpret_cast<uint4*>(sB2wl)[i] = __ldg(reinterpret_cast<const uint4*>(a.B2lo) + i);
    }
    if (tid < 16) sB2[tid] = a.b2p[tid];
    // chunks of A1 beyond the cond chunk are constant zero
    for (uint32_t i = tid + (uint32_t)(bg.npairs + 1) * 128; i < bg.a1_bytes / 16; i += BT_THREADS) {
        reinterpret_cast<uint4*>(sA1)[i] = make_uint4(0, 0, 0, 0);
        if (X3) reinterpret_cast<uint4*>(sA1l)[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)bg.tmem_cols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t tmem_d2_col = (uint32_t)bg.N1;
    const uint32_t idesc1 = X3 ? umma_idesc_f16(128, bg.N1) : umma_idesc_bf16(128, bg.N1), idesc2 = X3 ? umma_idesc_f16(128, 16) : umma_idesc_bf16(128, 16);
    const uint32_t lbo_b1 = (uint32_t)(g.fc / 8) * 128u;
    uint32_t phase = 0;
    const int py = tid >> 5, px = tid & 31;
    const uint32_t row_off = (uint32_t)(tid >> 3) * 128u + (uint32_t)(tid & 7) * 16u;   // this thread's row inside a K-chunk

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const DyncaTile t = dynca_tile_of(tile, a.tiles_x, a.tiles_y);
        const int gy = t.y0 + py, gx = t.x0 + px;
        const bool inimg = gy < H && gx < W;
        dynca_stage2_issue<NS, BT_THREADS>(g, a.x_in, a.xc, t, sStage);
        dynca_stage2_finish<NS, BT_THREADS>(g, sStage);
        // ---- perception -> A1 (this thread's row) ----
        {
            DyncaUp u = {};
            if (NS == 2 && inimg) u = dynca_up_of(g, t, gy, gx);
            for (int cp = 0; cp < bg.npairs; ++cp) {
                float f0[4] = {0.f, 0.f, 0.f, 0.f}, f1[4] = {0.f, 0.f, 0.f, 0.f};
                if (inimg) {
                    dynca_cell_percept2<NS>(g, sStage, u, 2 * cp, py, px, f0);
                    if (2 * cp + 1 < C) dynca_cell_percept2<NS>(g, sStage, u, 2 * cp + 1, py, px, f1);
                }
                uint4 v;
                if (X3) {
                    uint4 l;
                    split_f16x2(f0[0], f0[1], v.x, l.x); split_f16x2(f0[2], f0[3], v.y, l.y);
                    split_f16x2(f1[0], f1[1], v.z, l.z); split_f16x2(f1[2], f1[3], v.w, l.w);
                    *reinterpret_cast<uint4*>(sA1l + (uint32_t)cp * 2048u + row_off) = l;
                } else {
                    v.x = pack_bf16(f0[0], f0[1]); v.y = pack_bf16(f0[2], f0[3]);
                    v.z = pack_bf16(f1[0], f1[1]); v.w = pack_bf16(f1[2], f1[3]);
                }
                *reinterpret_cast<uint4*>(sA1 + (uint32_t)cp * 2048u + row_off) = v;
            }
            if (X3) {
                uint4 ch, cl;
                dynca_cond_chunk_x3(g, a.cond, t.b, gy, gx, inimg, ch, cl);
                *reinterpret_cast<uint4*>(sA1 + (uint32_t)bg.npairs * 2048u + row_off) = ch;
                *reinterpret_cast<uint4*>(sA1l + (uint32_t)bg.npairs * 2048u + row_off) = cl;
            } else {
                *reinterpret_cast<uint4*>(sA1 + (uint32_t)bg.npairs * 2048u + row_off) = dynca_cond_chunk(g, a.cond, t.b, gy, gx, inimg);
            }
        }
        fence_proxy_async();     // generic-proxy smem writes -> visible to the tensor-core (async) proxy
        tc_fence_before();
        __syncthreads();
        // ---- GEMM1 (X3: Ah.Bh + Al.Bh + Ah.Bl) ----
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sA1), b_addr = smem_u32(sB1);
            const uint32_t al_addr = smem_u32(sA1l), bl_addr = smem_u32(sB1l);
            for (int ks = 0; ks < bg.K1 / 16; ++ks) {
                const uint64_t da = umma_desc(a_addr + (uint32_t)ks * 2u * 2048u, 2048u, 128u);
                const uint64_t db = umma_desc(b_addr + (uint32_t)ks * 2u * lbo_b1, lbo_b1, 128u);
                umma_f16_ss(tmem_base, da, db, idesc1, ks > 0 ? 1u : 0u);
                if (X3) {
                    umma_f16_ss(tmem_base, umma_desc(al_addr + (uint32_t)ks * 2u * 2048u, 2048u, 128u), db, idesc1, 1u);
                    umma_f16_ss(tmem_base, da, umma_desc(bl_addr + (uint32_t)ks * 2u * lbo_b1, lbo_b1, 128u), idesc1, 1u);
                }
            }
            umma_commit(bar);
        }
        // while GEMM1 runs: fetch this cell's state (residual) and fire decision for epilogue 2
        float xin[16];
        float fire = 0.0f;
        const size_t off = (size_t)t.b * C * plane + (size_t)gy * W + gx;
        if (inimg) {
#pragma unroll
            for (int c = 0; c < 16; ++c) xin[c] = c < C ? __ldg(a.x_in + off + c * plane) : 0.0f;
            fire = dynca_fire(a.fm, t.b, gy, gx, H, W);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        __syncwarp();            // tcgen05.ld is .sync.aligned: reconverge after the single-thread issue branch
        tc_fence_after();
        // ---- epilogue 1: D1 -> relu -> bf16 -> A2 (overlays the stage area: all perception reads are done) ----
        for (int j0 = 0; j0 < bg.N1; j0 += 32) {
            uint32_t v[32];
            if (bg.N1 - j0 >= 32) {
                tmem_ld32(tmem_lane + (uint32_t)j0, v);
            } else {   // fc % 32 == 16: last 16 columns
                tmem_ld16(tmem_lane + (uint32_t)j0, v);
#pragma unroll
                for (int i = 16; i < 32; ++i) v[i] = 0u;
            }
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (j0 + q * 8 < bg.N1) {
                    uint4 o;
                    if (X3) {
                        uint4 l;
                        // D1 carries the weight scale: h = relu(D1) * 2^-8 (exact)
                        split_f16x2(fmaxf(__uint_as_float(v[q * 8 + 0]), 0.f) * NCA_X3_WINV, fmaxf(__uint_as_float(v[q * 8 + 1]), 0.f) * NCA_X3_WINV, o.x, l.x);
                        split_f16x2(fmaxf(__uint_as_float(v[q * 8 + 2]), 0.f) * NCA_X3_WINV, fmaxf(__uint_as_float(v[q * 8 + 3]), 0.f) * NCA_X3_WINV, o.y, l.y);
                        split_f16x2(fmaxf(__uint_as_float(v[q * 8 + 4]), 0.f) * NCA_X3_WINV, fmaxf(__uint_as_float(v[q * 8 + 5]), 0.f) * NCA_X3_WINV, o.z, l.z);
                        split_f16x2(fmaxf(__uint_as_float(v[q * 8 + 6]), 0.f) * NCA_X3_WINV, fmaxf(__uint_as_float(v[q * 8 + 7]), 0.f) * NCA_X3_WINV, o.w, l.w);
                        *reinterpret_cast<uint4*>(sA2l + (uint32_t)(j0 / 8 + q) * 2048u + row_off) = l;
                    } else {
                        o.x = pack_bf16(fmaxf(__uint_as_float(v[q * 8 + 0]), 0.f), fmaxf(__uint_as_float(v[q * 8 + 1]), 0.f));
                        o.y = pack_bf16(fmaxf(__uint_as_float(v[q * 8 + 2]), 0.f), fmaxf(__uint_as_float(v[q * 8 + 3]), 0.f));
                        o.z = pack_bf16(fmaxf(__uint_as_float(v[q * 8 + 4]), 0.f), fmaxf(__uint_as_float(v[q * 8 + 5]), 0.f));
                        o.w = pack_bf16(fmaxf(__uint_as_float(v[q * 8 + 6]), 0.f), fmaxf(__uint_as_float(v[q * 8 + 7]), 0.f));
                    }
                    *reinterpret_cast<uint4*>(sA2 + (uint32_t)(j0 / 8 + q) * 2048u + row_off) = o;
                }
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        // ---- GEMM2 ----
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sA2), b_addr = smem_u32(sB2w);
            const uint32_t al_addr = smem_u32(sA2l), bl_addr = smem_u32(sB2wl);
            for (int ks = 0; ks < bg.N1 / 16; ++ks) {
                const uint64_t da = umma_desc(a_addr + (uint32_t)ks * 2u * 2048u, 2048u, 128u);
                const uint64_t db = umma_desc(b_addr + (uint32_t)ks * 2u * 256u, 256u, 128u);
                umma_f16_ss(tmem_base + tmem_d2_col, da, db, idesc2, ks > 0 ? 1u : 0u);
                if (X3) {
                    umma_f16_ss(tmem_base + tmem_d2_col, umma_desc(al_addr + (uint32_t)ks * 2u * 2048u, 2048u, 128u), db, idesc2, 1u);
                    umma_f16_ss(tmem_base + tmem_d2_col, da, umma_desc(bl_addr + (uint32_t)ks * 2u * 256u, 256u, 128u), idesc2, 1u);
                }
            }
            umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        __syncwarp();            // tcgen05.ld is .sync.aligned: reconverge after the single-thread issue branch
        tc_fence_after();
        // ---- epilogue 2: y = D2 + b2 ; x' = x + y * fire ----
        {
            uint32_t v[16];
            tmem_ld16(tmem_lane + tmem_d2_col, v);
            tmem_ld_wait();
            if (inimg) {
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    if (c < C) a.x_out[off + c * plane] = xin[c] + (__uint_as_float(v[c]) * (X3 ? NCA_X3_WINV : 1.0f) + sB2[c]) * fire;
            }
        }
        tc_fence_before();
        __syncthreads();   // stage area / A1 are rewritten by the next tile; TMEM reads are complete
    }
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)bg.tmem_cols);
}

// =====================================================================================================
// BPTT step on the tensor cores.  256 threads: thread (m = tid & 127, half = tid >> 7) works on cell m.
//   S1: D1 = A1 . W1^T (recompute a)          D3 = Gy . W2   (g_h)                       [128 cells x fc]
//   E1: h = relu(D1), g_a = D3 * [D1 > 0]  ->  bf16 operands sH, sGa
//   S2: D4 += H^T . Gy   (gW2, [fc x 16])     D5 += G_a^T . Z  (gW1 | gb1, [fc x K1])    accumulated in TMEM over
//       D6  = G_a . W1   (g_z, [128 cells x K1])                                          all tiles of the CTA
//   E2: D6 -> s0 * g_z as fp32 [k'][cell] -> transposed perception -> red.add into g_t
// The weight-gradient operands are the SAME shared-memory tiles viewed MN-major (cells become K):
//   tile (cell m, col j) at (j/8)*2048 + (m/8)*128 + (m%8)*16 + (j%8)*2  ==  K-major  [m][j]  (LBO 2048, SBO 128)
//                                                                      ==  MN-major [j][m]  (SBO 2048, LBO 128)
// =====================================================================================================
#define BB_THREADS 256
#define BB_TMEM_D1 0u
#define BB_TMEM_D3 128u
#define BB_TMEM_D4 256u
#define BB_TMEM_D5 272u
#define BB_TMEM_D6 352u


// B1t [N=K1][K=fc] (dgrad1) and B2d [N=fc][K=16] (dgrad2) operand images
__global__ void dynca_bf16_prep_bwd_kernel(DyncaGeom g, Bf16Geom bg, const float* __restrict__ w1, const float* __restrict__ b1,
                                           const float* __restrict__ w2, __nv_bfloat16* __restrict__ B1t,
                                           __nv_bfloat16* __restrict__ B2d, __nv_bfloat16* __restrict__ B2dlo) {
    const int n1 = bg.K1 * g.fc, n2 = g.fc * 16;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += gridDim.x * blockDim.x) {
        if (i < n1) {
            const int kp = i / g.fc, j = i % g.fc;
            const int kc = kp >> 3, s = kp & 7;
            float v = 0.0f;
            if (kc < bg.npairs) {
                const int c = 2 * kc + (s >> 2), f = s & 3;
                if (c < g.C) v = w1[j * g.P + f * g.C + c];
            }   // cond / bias columns of g_z are never used: leave zero
            const size_t off = (size_t)(j >> 3) * (bg.K1 / 8) * 64 + (size_t)kc * 64 + s * 8 + (j & 7);
            if (B1t) B1t[off] = __float2bfloat16_rn(v);
        } else {
            const int e = i - n1, j = e / 16, c = e % 16;
            const float v = c < g.C ? w2[c * g.fc + j] : 0.0f;
            const size_t off = (size_t)(c >> 3) * (g.fc / 8) * 64 + (size_t)(j >> 3) * 64 + (j & 7) * 8 + (c & 7);
            if (B2dlo) {
                const __half h = __float2half_rn(v * NCA_X3_WSCALE);
                reinterpret_cast<__half*>(B2d)[off] = h;
                reinterpret_cast<__half*>(B2dlo)[off] = __float2half_rn(v * NCA_X3_WSCALE - __half2float(h));
            } else {
                B2d[off] = __float2bfloat16_rn(v);
            }
        }
    }
}

// X3: max |dL/dx_{t+1}| (tap included) of one BPTT step -> *out (non-negative floats order like their bit patterns)
__global__ void dynca_absmax_kernel(int B, int C, size_t plane, const float* __restrict__ g_next, const float* __restrict__ g_tap, int tap_c,
                                    float tap_scale, float* __restrict__ out) {
    const size_t n = (size_t)B * C * plane;
    float m = 0.0f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t pix = i % plane;
        const int c = (int)((i / plane) % C), b = (int)(i / (plane * C));
        m = fmaxf(m, fabsf(dynca_gnext(g_next, g_tap, tap_c, tap_scale, C, b, c, pix, plane)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));
}

struct DyncaBf16BwdArgs {
    DyncaGeom g;
    Bf16Geom bg;
    const float* x_in; const float* xc; const float* g_next; const float* g_tap; int tap_c; float tap_scale;
    float* g_out; const float* cond;
    const __nv_bfloat16* B1; const __nv_bfloat16* B1t; const __nv_bfloat16* B2d;
    const __nv_bfloat16* B1lo; const __nv_bfloat16* B2dlo;      // X3: lo images (B1t is not used: D6 reads B1 through an MN-major view)
    const float* gmax;                                          // X3: max |dL/dx_{t+1}| of this launch (device scalar) -> gradient-operand scale
    float* gW1p; float* gW2p; float* gb2p;      // fp32 accumulators, padded fp32-path layout (red.add)
    FireMask fm;
    int tiles_x, tiles_y, n_tiles;
};

static inline size_t dynca_bf16_bwd_smem_bytes(const DyncaGeom& g, const Bf16Geom& bg, bool x3) {
    // X3: hi + lo image of every operand; the fp32 stage overlays H | Ga (no prefetch of the next tile)
    if (x3) return 1024 + 128 + 2 * ((size_t)bg.a1_bytes + 4096 + 2 * 32768 + bg.b1_bytes + (size_t)2 * (g.fc / 8) * 128);
    return 1024 + 128 + bg.a1_bytes + 4096 /*Gy*/ + 2 * 32768 /*H, Ga*/ + 2 * bg.b1_bytes + (size_t)2 * (g.fc / 8) * 128 +
           (size_t)dynca_stage2_floats(g) * 4;
}

template <int NS, bool X3>
__global__ void __launch_bounds__(BB_THREADS, 1) dynca_bwd_bf16_kernel(const DyncaBf16BwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const DyncaGeom& g = a.g;
    const Bf16Geom& bg = a.bg;
    uint8_t* base = smem_raw;   // 1024-byte aligned by declaration; no integer round-trip, so accesses stay LDS/STS
    uint64_t* bar = reinterpret_cast<uint64_t*>(base);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + 8);
    // X3: every operand region holds the hi image followed by the lo image
    constexpr uint32_t XM = X3 ? 2u : 1u;
    const uint32_t b2d_bytes = (uint32_t)(2 * (g.fc / 8) * 128);
    uint8_t* sA1 = base + 128;
    uint8_t* sA1l = sA1 + bg.a1_bytes;
    uint8_t* sGy = sA1 + XM * bg.a1_bytes;
    uint8_t* sGyl = sGy + 4096;
    uint8_t* sH = sGy + XM * 4096;
    uint8_t* sHl = sH + 32768;
    uint8_t* sGa = sH + XM * 32768;
    uint8_t* sGal = sGa + 32768;
    uint8_t* sB1 = sGa + XM * 32768;
    uint8_t* sB1t = sB1 + bg.b1_bytes;               // X3: the lo image of B1 lives here instead
    uint8_t* sB1l = sB1t;
    uint8_t* sB2d = sB1t + bg.b1_bytes;
    uint8_t* sB2dl = sB2d + b2d_bytes;
    float* sStage = X3 ? reinterpret_cast<float*>(sH) : reinterpret_cast<float*>(sB2d + b2d_bytes);
    float* sGz = reinterpret_cast<float*>(sH);       // fp32 zero-padded planes [8*npairs][4][SP2_S], overlay H | Ga after S2
    float* sScr = sGz + 8 * bg.npairs * SP2_PLANE;   // coarse planes + bilinear weights of the scatter (NS == 2), same overlay
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = tid & 127, half = tid >> 7;
    const int C = g.C, H = g.H, W = g.W, fc = g.fc;
    const size_t plane = (size_t)H * W;

    for (uint32_t i = tid; i < bg.b1_bytes / 16; i += BB_THREADS) {
        reinterpret_cast<uint4*>(sB1)[i] = __ldg(reinterpret_cast<const uint4*>(a.B1) + i);
        reinterpret_cast<uint4*>(sB1t)[i] = __ldg(reinterpret_cast<const uint4*>(X3 ? a.B1lo : a.B1t) + i);
    }
    for (uint32_t i = tid; i < b2d_bytes / 16; i += BB_THREADS) {
        reinterpret_cast<uint4*>(sB2d)[i] = __ldg(reinterpret_cast<const uint4*>(a.B2d) + i);
        if (X3) reinterpret_cast<uint4*>(sB2dl)[i] = __ldg(reinterpret_cast<const uint4*>(a.B2dlo) + i);
    }
    // zero everything the MMAs read but the tile loop never writes: A1 chunks past the cond chunk, H / Ga chunks
    // past fc/8 (rows fc..127 of the M=128 weight-gradient operands)
    for (uint32_t i = tid; i < XM * bg.a1_bytes / 16; i += BB_THREADS) reinterpret_cast<uint4*>(sA1)[i] = make_uint4(0, 0, 0, 0);
    for (uint32_t i = tid; i < (XM * 2u * 32768u) / 16; i += BB_THREADS) reinterpret_cast<uint4*>(sH)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(tmem_slot, 512u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t idesc_fc = X3 ? umma_idesc_f16(128, fc) : umma_idesc_bf16(128, fc), idesc_k1 = X3 ? umma_idesc_f16(128, bg.K1) : umma_idesc_bf16(128, bg.K1);
    const uint32_t idesc_w2 = X3 ? umma_idesc_f16_mn(128, 16) : umma_idesc_bf16_mn(128, 16), idesc_w1 = X3 ? umma_idesc_f16_mn(128, bg.K1) : umma_idesc_bf16_mn(128, bg.K1);
    // X3: gradient operands are scaled by sg (a power of two) so that they sit in fp16's range; undone exactly in E2 / the flush
    const float sg = X3 ? nca_x3_gscale(__ldg(a.gmax)) : 1.0f, sg_inv = 1.0f / sg;
    const uint32_t lbo_fc = (uint32_t)(fc / 8) * 128u, lbo_k1 = (uint32_t)(bg.K1 / 8) * 128u;
    uint32_t phase = 0;
    const int py = m >> 5, px = m & 31;
    const uint32_t row_off = (uint32_t)(m >> 3) * 128u + (uint32_t)(m & 7) * 16u;
    float b2acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) b2acc[i] = 0.0f;
    bool first = true;

    if (!X3 && (int)blockIdx.x < a.n_tiles)
        dynca_stage2_issue<NS, BB_THREADS>(g, a.x_in, a.xc, dynca_tile_of(blockIdx.x, a.tiles_x, a.tiles_y), sStage);
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const DyncaTile t = dynca_tile_of(tile, a.tiles_x, a.tiles_y);
        const int gy = t.y0 + py, gx = t.x0 + px;
        const bool inimg = gy < H && gx < W;
        if (X3) dynca_stage2_issue<NS, BB_THREADS>(g, a.x_in, a.xc, t, sStage);      // the stage overlays H | Ga: no prefetch
        dynca_stage2_finish<NS, BB_THREADS>(g, sStage);
        // ---- recompute perception -> A1 ----
        {
            DyncaUp u = {};
            if (NS == 2 && inimg) u = dynca_up_of(g, t, gy, gx);
            for (int cp = half; cp < bg.npairs; cp += 2) {
                float f0[4] = {0.f, 0.f, 0.f, 0.f}, f1[4] = {0.f, 0.f, 0.f, 0.f};
                if (inimg) {
                    dynca_cell_percept2<NS>(g, sStage, u, 2 * cp, py, px, f0);
                    if (2 * cp + 1 < C) dynca_cell_percept2<NS>(g, sStage, u, 2 * cp + 1, py, px, f1);
                }
                uint4 v;
                if (X3) {
                    uint4 l;
                    split_f16x2(f0[0], f0[1], v.x, l.x); split_f16x2(f0[2], f0[3], v.y, l.y);
                    split_f16x2(f1[0], f1[1], v.z, l.z); split_f16x2(f1[2], f1[3], v.w, l.w);
                    *reinterpret_cast<uint4*>(sA1l + (uint32_t)cp * 2048u + row_off) = l;
                } else {
                    v.x = pack_bf16(f0[0], f0[1]); v.y = pack_bf16(f0[2], f0[3]);
                    v.z = pack_bf16(f1[0], f1[1]); v.w = pack_bf16(f1[2], f1[3]);
                }
                *reinterpret_cast<uint4*>(sA1 + (uint32_t)cp * 2048u + row_off) = v;
            }
            if (half == 0) {
                if (X3) {
                    uint4 ch, cl;
                    dynca_cond_chunk_x3(g, a.cond, t.b, gy, gx, inimg, ch, cl);
                    *reinterpret_cast<uint4*>(sA1 + (uint32_t)bg.npairs * 2048u + row_off) = ch;
                    *reinterpret_cast<uint4*>(sA1l + (uint32_t)bg.npairs * 2048u + row_off) = cl;
                } else {
                    *reinterpret_cast<uint4*>(sA1 + (uint32_t)bg.npairs * 2048u + row_off) = dynca_cond_chunk(g, a.cond, t.b, gy, gx, inimg);
                }
            }
        }
        // ---- g_y = fire * g_{t+1}: this thread's 8 channels of its cell ----
        {
            float gyv[8];
            const float fire = inimg ? dynca_fire(a.fm, t.b, gy, gx, H, W) : 0.0f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = 8 * half + i;
                gyv[i] = (inimg && c < C) ? fire * dynca_gnext(a.g_next, a.g_tap, a.tap_c, a.tap_scale, C, t.b, c, (size_t)gy * W + gx, plane) : 0.0f;
                b2acc[i] += gyv[i];
            }
            uint4 v;
            if (X3) {
                uint4 l;
                split_f16x2(gyv[0] * sg, gyv[1] * sg, v.x, l.x); split_f16x2(gyv[2] * sg, gyv[3] * sg, v.y, l.y);
                split_f16x2(gyv[4] * sg, gyv[5] * sg, v.z, l.z); split_f16x2(gyv[6] * sg, gyv[7] * sg, v.w, l.w);
                *reinterpret_cast<uint4*>(sGyl + (uint32_t)half * 2048u + row_off) = l;
            } else {
                v.x = pack_bf16(gyv[0], gyv[1]); v.y = pack_bf16(gyv[2], gyv[3]); v.z = pack_bf16(gyv[4], gyv[5]); v.w = pack_bf16(gyv[6], gyv[7]);
            }
            *reinterpret_cast<uint4*>(sGy + (uint32_t)half * 2048u + row_off) = v;
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        // every thread is done reading the stage area: start copying the next tile of this CTA behind the MMAs
        if (!X3 && tile + (int)gridDim.x < a.n_tiles)
            dynca_stage2_issue<NS, BB_THREADS>(g, a.x_in, a.xc, dynca_tile_of(tile + gridDim.x, a.tiles_x, a.tiles_y), sStage);
        // ---- S1: D1 = A1 . W1^T ; D3 = Gy . W2  (X3: Ah.Bh + Al.Bh + Ah.Bl) ----
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sA1), b_addr = smem_u32(sB1);
            for (int ks = 0; ks < bg.K1 / 16; ++ks) {
                const uint64_t da = umma_desc(a_addr + (uint32_t)ks * 4096u, 2048u, 128u);
                const uint64_t db = umma_desc(b_addr + (uint32_t)ks * 2u * lbo_fc, lbo_fc, 128u);
                umma_f16_ss(tmem_base + BB_TMEM_D1, da, db, idesc_fc, ks > 0 ? 1u : 0u);
                if (X3) {
                    umma_f16_ss(tmem_base + BB_TMEM_D1, umma_desc(smem_u32(sA1l) + (uint32_t)ks * 4096u, 2048u, 128u), db, idesc_fc, 1u);
                    umma_f16_ss(tmem_base + BB_TMEM_D1, da, umma_desc(smem_u32(sB1l) + (uint32_t)ks * 2u * lbo_fc, lbo_fc, 128u), idesc_fc, 1u);
                }
            }
            {
                const uint64_t dg = umma_desc(smem_u32(sGy), 2048u, 128u), dw = umma_desc(smem_u32(sB2d), lbo_fc, 128u);
                umma_f16_ss(tmem_base + BB_TMEM_D3, dg, dw, idesc_fc, 0u);
                if (X3) {
                    umma_f16_ss(tmem_base + BB_TMEM_D3, umma_desc(smem_u32(sGyl), 2048u, 128u), dw, idesc_fc, 1u);
                    umma_f16_ss(tmem_base + BB_TMEM_D3, dg, umma_desc(smem_u32(sB2dl), lbo_fc, 128u), idesc_fc, 1u);
                }
            }
            umma_commit(bar);
        }
        if (X3 && fc < 128) {      // the stage clobbered the zero tails (hidden rows fc..127) of H / Ga: restore them before E1 / S2
            for (uint32_t i = tid; i < (uint32_t)(16 - fc / 8) * 128u; i += BB_THREADS) {
                reinterpret_cast<uint4*>(sH + (uint32_t)(fc / 8) * 2048u)[i] = make_uint4(0, 0, 0, 0);
                reinterpret_cast<uint4*>(sHl + (uint32_t)(fc / 8) * 2048u)[i] = make_uint4(0, 0, 0, 0);
                reinterpret_cast<uint4*>(sGa + (uint32_t)(fc / 8) * 2048u)[i] = make_uint4(0, 0, 0, 0);
                reinterpret_cast<uint4*>(sGal + (uint32_t)(fc / 8) * 2048u)[i] = make_uint4(0, 0, 0, 0);
            }
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        __syncwarp();
        tc_fence_after();
        // ---- E1: h, g_a -> bf16 operands; half -> columns [64*half, 64*half + 64) ----
#pragma unroll 1
        for (int blk = 0; blk < 2; ++blk) {
            const int j0 = 64 * half + 32 * blk;
            if (j0 < fc) {     // warp-uniform
                uint32_t av[32], gv[32];
                tmem_ld32(tmem_lane + BB_TMEM_D1 + (uint32_t)j0, av);
                tmem_ld32(tmem_lane + BB_TMEM_D3 + (uint32_t)j0, gv);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float hh[8], ga[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float av_ = __uint_as_float(av[q * 8 + i]);
                        hh[i] = fmaxf(av_, 0.0f);
                        ga[i] = av_ > 0.0f ? __uint_as_float(gv[q * 8 + i]) : 0.0f;
                        if (X3) { hh[i] *= NCA_X3_WINV; ga[i] *= NCA_X3_WINV; }      // D1, D3 carry the weight scale (ga keeps sg)
                    }
                    uint4 o;
                    if (X3) {
                        uint4 l;
                        split_f16x2(hh[0], hh[1], o.x, l.x); split_f16x2(hh[2], hh[3], o.y, l.y);
                        split_f16x2(hh[4], hh[5], o.z, l.z); split_f16x2(hh[6], hh[7], o.w, l.w);
                        *reinterpret_cast<uint4*>(sH + (uint32_t)(j0 / 8 + q) * 2048u + row_off) = o;
                        *reinterpret_cast<uint4*>(sHl + (uint32_t)(j0 / 8 + q) * 2048u + row_off) = l;
                        split_f16x2(ga[0], ga[1], o.x, l.x); split_f16x2(ga[2], ga[3], o.y, l.y);
                        split_f16x2(ga[4], ga[5], o.z, l.z); split_f16x2(ga[6], ga[7], o.w, l.w);
                        *reinterpret_cast<uint4*>(sGa + (uint32_t)(j0 / 8 + q) * 2048u + row_off) = o;
                        *reinterpret_cast<uint4*>(sGal + (uint32_t)(j0 / 8 + q) * 2048u + row_off) = l;
                        continue;
                    }
                    o.x = pack_bf16(hh[0], hh[1]); o.y = pack_bf16(hh[2], hh[3]); o.z = pack_bf16(hh[4], hh[5]); o.w = pack_bf16(hh[6], hh[7]);
                    *reinterpret_cast<uint4*>(sH + (uint32_t)(j0 / 8 + q) * 2048u + row_off) = o;
                    o.x = pack_bf16(ga[0], ga[1]); o.y = pack_bf16(ga[2], ga[3]); o.z = pack_bf16(ga[4], ga[5]); o.w = pack_bf16(ga[6], ga[7]);
                    *reinterpret_cast<uint4*>(sGa + (uint32_t)(j0 / 8 + q) * 2048u + row_off) = o;
                }
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        // ---- S2: weight gradients (accumulated in TMEM across tiles) and g_z ----
        if (tid == 0) {
            tc_fence_after();
            const uint32_t h_addr = smem_u32(sH), ga_addr = smem_u32(sGa), gy_addr = smem_u32(sGy), z_addr = smem_u32(sA1);
            const uint32_t hl_addr = smem_u32(sHl), gal_addr = smem_u32(sGal), gyl_addr = smem_u32(sGyl), zl_addr = smem_u32(sA1l);
            for (int ks = 0; ks < 8; ++ks) {   // 16 cells per instruction
                const uint32_t acc = (first && ks == 0) ? 0u : 1u;
                const uint64_t dh = umma_desc(h_addr + (uint32_t)ks * 256u, 128u, 2048u), dgy = umma_desc(gy_addr + (uint32_t)ks * 256u, 128u, 2048u);
                const uint64_t dga = umma_desc(ga_addr + (uint32_t)ks * 256u, 128u, 2048u), dz = umma_desc(z_addr + (uint32_t)ks * 256u, 128u, 2048u);
                umma_f16_ss(tmem_base + BB_TMEM_D4, dh, dgy, idesc_w2, acc);
                umma_f16_ss(tmem_base + BB_TMEM_D5, dga, dz, idesc_w1, acc);
                if (X3) {
                    umma_f16_ss(tmem_base + BB_TMEM_D4, umma_desc(hl_addr + (uint32_t)ks * 256u, 128u, 2048u), dgy, idesc_w2, 1u);
                    umma_f16_ss(tmem_base + BB_TMEM_D4, dh, umma_desc(gyl_addr + (uint32_t)ks * 256u, 128u, 2048u), idesc_w2, 1u);
                    umma_f16_ss(tmem_base + BB_TMEM_D5, umma_desc(gal_addr + (uint32_t)ks * 256u, 128u, 2048u), dz, idesc_w1, 1u);
                    umma_f16_ss(tmem_base + BB_TMEM_D5, dga, umma_desc(zl_addr + (uint32_t)ks * 256u, 128u, 2048u), idesc_w1, 1u);
                }
            }
            if (X3) {
                // D6 = Ga . W1 with W1 read from the forward image B1 through an MN-major view ([N = k'][K = hidden]: LBO 128 = next 8
                // hidden units, SBO lbo_fc = next 8 k'); the cond / bias columns of g_z are computed and never read
                const uint32_t idesc_gz = idesc_k1 | (1u << 16);
                for (int ks = 0; ks < fc / 16; ++ks) {
                    const uint64_t dga = umma_desc(ga_addr + (uint32_t)ks * 4096u, 2048u, 128u);
                    const uint64_t dw = umma_desc(smem_u32(sB1) + (uint32_t)ks * 256u, 128u, lbo_fc);
                    umma_f16_ss(tmem_base + BB_TMEM_D6, dga, dw, idesc_gz, ks > 0 ? 1u : 0u);
                    umma_f16_ss(tmem_base + BB_TMEM_D6, umma_desc(gal_addr + (uint32_t)ks * 4096u, 2048u, 128u), dw, idesc_gz, 1u);
                    umma_f16_ss(tmem_base + BB_TMEM_D6, dga, umma_desc(smem_u32(sB1l) + (uint32_t)ks * 256u, 128u, lbo_fc), idesc_gz, 1u);
                }
            } else {
                const uint32_t bt_addr = smem_u32(sB1t);
                for (int ks = 0; ks < fc / 16; ++ks)
                    umma_f16_ss(tmem_base + BB_TMEM_D6, umma_desc(ga_addr + (uint32_t)ks * 4096u, 2048u, 128u),
                                umma_desc(bt_addr + (uint32_t)ks * 2u * lbo_k1, lbo_k1, 128u), idesc_k1, ks > 0 ? 1u : 0u);
            }
            umma_commit(bar);
        }
        first = false;
        mbar_wait(bar, phase);
        phase ^= 1u;
        __syncwarp();
        tc_fence_after();
        // ---- E2: g_z -> fp32 [k'][cell] (overlays H | Ga: every MMA that read them has completed) ----
        {
            const int k0 = 32 * half;
            if (k0 < 8 * bg.npairs) {
                uint32_t v[32];
                tmem_ld32(tmem_lane + BB_TMEM_D6 + (uint32_t)k0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (k0 + i < 8 * bg.npairs) sGz[(k0 + i) * SP2_PLANE + py * SP2_S + px + 1] = __uint_as_float(v[i]) * (X3 ? g.s0 * NCA_X3_WINV * sg_inv : g.s0);
            }
            // zero columns 0 and 33..39 of every plane row
            for (int i = tid; i < 8 * bg.npairs * DT_TH * 8; i += BB_THREADS) {
                const int q = i & 7, r = i >> 3;
                sGz[r * SP2_S + (q == 0 ? 0 : 32 + q)] = 0.0f;
            }
        }
        tc_fence_before();
        __syncthreads();
        dynca_scatter_tile_v2<NS, BB_THREADS>(g, t, sGz, sScr, reinterpret_cast<float*>(sGy), a.g_out, a.g_next, a.g_tap, a.tap_c, a.tap_scale);
        __syncthreads();
        // H / Ga rows past fc and the zero tail were clobbered by sGz: restore the zeros the next S2 relies on
        // (X3: done after the stage of the next tile has been consumed, see above)
        if (!X3 && fc < 128)
            for (uint32_t i = tid; i < (uint32_t)(16 - fc / 8) * 128u; i += BB_THREADS) {
                reinterpret_cast<uint4*>(sH + (uint32_t)(fc / 8) * 2048u)[i] = make_uint4(0, 0, 0, 0);
                reinterpret_cast<uint4*>(sGa + (uint32_t)(fc / 8) * 2048u)[i] = make_uint4(0, 0, 0, 0);
            }
    }
    // ---- flush: D4 [fc x 16] -> gW2p[j][c];  D5 [fc x K1] -> gW1p[k][j] (k' -> reference k);  gb2 ----
    {
        const int j = (warp & 3) * 32 + lane;
        uint32_t v[32];
        if (half == 0) {
            tmem_ld16(tmem_lane + BB_TMEM_D4, v);
            tmem_ld_wait();
            if (j < fc)
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    if (c < C) atomicAdd(a.gW2p + j * g.CP + c, __uint_as_float(v[c]) * sg_inv);
        }
        for (int k0 = 32 * half; k0 < bg.K1; k0 += 64) {
            if (bg.K1 - k0 >= 32) tmem_ld32(tmem_lane + BB_TMEM_D5 + (uint32_t)k0, v);
            else tmem_ld16(tmem_lane + BB_TMEM_D5 + (uint32_t)k0, v);
            tmem_ld_wait();
            if (j < fc) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int kp = k0 + i;
                    if (kp >= bg.K1 || (i >= 16 && bg.K1 - k0 < 32)) continue;
                    const int kc = kp >> 3, s = kp & 7;
                    int k = -1;
                    if (kc < bg.npairs) { const int c = 2 * kc + (s >> 2); if (c < C) k = (s & 3) * C + c; }
                    else if (kc == bg.npairs) { const int src = dynca_cond_slot_src(g.cc, s); if (src >= 0) k = 4 * C + src; else if (src == -2) k = g.P; }
                    if (k >= 0) atomicAdd(a.gW1p + k * g.FCpad + j, __uint_as_float(v[i]) * sg_inv);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float s = b2acc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0 && 8 * half + i < C) atomicAdd(a.gb2p + 8 * half + i, s);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512u);
}

// ---- host launchers -----------------------------------------------------------------------------------
static int bf16_num_sms() { return nca_sm_count(); }

// forward operand images: B1 | B2 | b2 (64 B) [| B1 lo | B2 lo with x3]
size_t dynca_bf16_weight_bytes(const DyncaGeom& g, bool x3) {
    Bf16Geom bg;
    if (dynca_bf16_geom(g, &bg)) return 0;
    return nca_align_up((size_t)(x3 ? 2 : 1) * (bg.b1_bytes + bg.b2_bytes) + 64, 256);
}

int dynca_bf16_prep_weights(const DyncaGeom& g, const NcaDyncaWeights* w, void* ws, cudaStream_t s, bool x3) {
    Bf16Geom bg;
    int rc = dynca_bf16_geom(g, &bg);
    if (rc) return rc;
    __nv_bfloat16* B1 = (__nv_bfloat16*)ws;
    __nv_bfloat16* B2 = (__nv_bfloat16*)((uint8_t*)ws + bg.b1_bytes);
    float* b2p = (float*)((uint8_t*)ws + bg.b1_bytes + bg.b2_bytes);
    __nv_bfloat16* B1lo = x3 ? (__nv_bfloat16*)((uint8_t*)ws + bg.b1_bytes + bg.b2_bytes + 64) : nullptr;
    __nv_bfloat16* B2lo = x3 ? (__nv_bfloat16*)((uint8_t*)ws + 2 * bg.b1_bytes + bg.b2_bytes + 64) : nullptr;
    dynca_bf16_prep_kernel<<<32, 256, 0, s>>>(g, bg, w->w1, w->b1, w->w2, w->b2, B1, B2, b2p, B1lo, B2lo);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int dynca_bf16_coarsen(const DyncaGeom& g, const float* x, float* xc, cudaStream_t s) {
    if (g.ns != 2) return NCA_OK;
    const size_t n = (size_t)g.B * g.C * (g.H / 2) * (g.W / 2);
    int grid = (int)((n + 255) / 256 < (size_t)bf16_num_sms() * 8 ? (n + 255) / 256 : (size_t)bf16_num_sms() * 8);
    dynca_coarsen_kernel<<<grid, 256, 0, s>>>(g.B * g.C, g.H, g.W, x, xc);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

bool dynca_bf16_supported(const DyncaGeom& g, bool x3) {
    Bf16Geom bg;
    if (dynca_bf16_geom(g, &bg)) return false;
    return dynca_bf16_smem_bytes(g, bg, x3) <= 227 * 1024;
}

int dynca_bf16_forward_step(const DyncaGeom& g, const void* ws, float* xc, const float* x_in, float* x_out, const float* cond,
                            const FireMask& fm, cudaStream_t s, bool x3) {
    DyncaBf16Args a;
    int rc = dynca_bf16_geom(g, &a.bg);
    if (rc) return rc;
    rc = dynca_bf16_coarsen(g, x_in, xc, s);   // (per step: this variant does not produce the next coarse state itself)
    if (rc) return rc;
    a.g = g; a.x_in = x_in; a.xc = xc; a.x_out = x_out; a.cond = cond;
    a.B1 = (const __nv_bfloat16*)ws;
    a.B2 = (const __nv_bfloat16*)((const uint8_t*)ws + a.bg.b1_bytes);
    a.b2p = (const float*)((const uint8_t*)ws + a.bg.b1_bytes + a.bg.b2_bytes);
    a.B1lo = x3 ? (const __nv_bfloat16*)((const uint8_t*)ws + a.bg.b1_bytes + a.bg.b2_bytes + 64) : nullptr;
    a.B2lo = x3 ? (const __nv_bfloat16*)((const uint8_t*)ws + 2 * a.bg.b1_bytes + a.bg.b2_bytes + 64) : nullptr;
    a.fm = fm;
    a.tiles_x = (g.W + DT_TW - 1) / DT_TW; a.tiles_y = (g.H + DT_TH - 1) / DT_TH; a.n_tiles = g.B * a.tiles_x * a.tiles_y;
    const size_t smem = dynca_bf16_smem_bytes(g, a.bg, x3);
    if (smem > 227 * 1024) { nca_set_error("shared memory need %zu B exceeds 227 KB", smem); return NCA_ERR_UNSUPPORTED; }
    // resident CTAs per SM: limited by shared memory and by TMEM columns (512 per SM)
    int occ = (int)((227 * 1024) / (smem + 1024));
    if (occ > 512 / a.bg.tmem_cols) occ = 512 / a.bg.tmem_cols;
    if (occ > 4) occ = 4;
    if (occ < 1) occ = 1;
    int grid = bf16_num_sms() * occ;
    if (grid > a.n_tiles) grid = a.n_tiles;
#define BT_LAUNCH(NS_, X3_)                                                                                                       \
    do {                                                                                                                          \
        NCA_CUDA_OK(cudaFuncSetAttribute(dynca_fwd_bf16_kernel<NS_, X3_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        dynca_fwd_bf16_kernel<NS_, X3_><<<grid, BT_THREADS, smem, s>>>(a);                                                        \
    } while (0)
    if (g.ns == 2) { if (x3) BT_LAUNCH(2, true); else BT_LAUNCH(2, false); }
    else { if (x3) BT_LAUNCH(1, true); else BT_LAUNCH(1, false); }
    NCA_LAUNCH_OK();
    return NCA_OK;
}

// BPTT operand images: B1 | B1t | B2d [| B2d lo with x3; B1t's slot then holds the lo image of B1]
size_t dynca_bf16_bwd_weight_bytes(const DyncaGeom& g, bool x3) {
    Bf16Geom bg;
    if (dynca_bf16_geom(g, &bg)) return 0;
    return nca_align_up((size_t)2 * bg.b1_bytes + (size_t)(x3 ? 2 : 1) * 2 * (g.fc / 8) * 128 + 64, 256);
}
bool dynca_bf16_bwd_supported(const DyncaGeom& g, bool x3) {
    Bf16Geom bg;
    if (g.fc % 32 != 0 || g.fc > 128 || g.cc + 2 > 8) return false;
    if (dynca_bf16_geom(g, &bg)) return false;
    return bg.K1 <= 80 && dynca_bf16_bwd_smem_bytes(g, bg, x3) <= 227 * 1024;
}

int dynca_bf16_prep_bwd_weights(const DyncaGeom& g, const NcaDyncaWeights* w, void* ws, cudaStream_t s, bool x3) {
    Bf16Geom bg;
    int rc = dynca_bf16_geom(g, &bg);
    if (rc) return rc;
    const size_t b2d_bytes = (size_t)2 * (g.fc / 8) * 128;
    __nv_bfloat16* B1 = (__nv_bfloat16*)ws;
    __nv_bfloat16* B1t = (__nv_bfloat16*)((uint8_t*)ws + bg.b1_bytes);
    __nv_bfloat16* B2d = (__nv_bfloat16*)((uint8_t*)ws + 2 * bg.b1_bytes);
    __nv_bfloat16* B2dlo = x3 ? (__nv_bfloat16*)((uint8_t*)ws + 2 * bg.b1_bytes + b2d_bytes) : nullptr;
    dynca_bf16_prep_bwd_kernel<<<32, 256, 0, s>>>(g, bg, w->w1, w->b1, w->w2, x3 ? nullptr : B1t, B2d, B2dlo);
    NCA_LAUNCH_OK();
    dynca_bf16_prep_b1_kernel<<<32, 256, 0, s>>>(g, bg, w->w1, w->b1, B1, x3 ? B1t : nullptr);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int dynca_bf16_backward_step(const DyncaGeom& g, const void* ws, float* xc, const float* xc_ready, float* wsG, const float* x_in, const float* g_next,
                             const float* g_tap, int tap_c, float tap_scale, float* g_out, const float* cond,
                             const FireMask& fm, cudaStream_t s, bool x3) {
    DyncaBf16BwdArgs a;
    int rc = dynca_bf16_geom(g, &a.bg);
    if (rc) return rc;
    if (xc_ready == nullptr) {
        rc = dynca_bf16_coarsen(g, x_in, xc, s);
        if (rc) return rc;
    }
    a.g = g; a.x_in = x_in; a.xc = xc_ready ? xc_ready : xc; a.g_next = g_next; a.g_tap = g_tap; a.tap_c = tap_c; a.tap_scale = tap_scale;
    a.g_out = g_out; a.cond = cond;
    a.B1 = (const __nv_bfloat16*)ws;
    a.B1t = (const __nv_bfloat16*)((const uint8_t*)ws + a.bg.b1_bytes);
    a.B2d = (const __nv_bfloat16*)((const uint8_t*)ws + 2 * a.bg.b1_bytes);
    a.B1lo = x3 ? a.B1t : nullptr;
    a.B2dlo = x3 ? (const __nv_bfloat16*)((const uint8_t*)ws + 2 * a.bg.b1_bytes + (size_t)2 * (g.fc / 8) * 128) : nullptr;
    a.gmax = nullptr;
    if (x3) {      // the 64-byte tail of the operand images holds the gradient scale of this launch
        float* gm = (float*)((uint8_t*)ws + 2 * a.bg.b1_bytes + (size_t)4 * (g.fc / 8) * 128);
        NCA_CUDA_OK(cudaMemsetAsync(gm, 0, sizeof(float), s));
        const size_t n = (size_t)g.B * g.C * g.H * g.W;
        const int grid = (int)((n + 255) / 256 < (size_t)bf16_num_sms() * 8 ? (n + 255) / 256 : (size_t)bf16_num_sms() * 8);
        dynca_absmax_kernel<<<grid, 256, 0, s>>>(g.B, g.C, (size_t)g.H * g.W, g_next, g_tap, tap_c, tap_scale, gm);
        NCA_LAUNCH_OK();
        a.gmax = gm;
    }
    a.gW1p = wsG; a.gW2p = a.gW1p + (size_t)g.Ppad * g.FCpad; a.gb2p = a.gW2p + (size_t)g.FCpad * g.CP;
    a.fm = fm;
    a.tiles_x = (g.W + DT_TW - 1) / DT_TW; a.tiles_y = (g.H + DT_TH - 1) / DT_TH; a.n_tiles = g.B * a.tiles_x * a.tiles_y;
    const size_t smem = dynca_bf16_bwd_smem_bytes(g, a.bg, x3);
    int grid = bf16_num_sms();
    if (grid > a.n_tiles) grid = a.n_tiles;
#define BB_LAUNCH(NS_, X3_)                                                                                                       \
    do {                                                                                                                          \
        NCA_CUDA_OK(cudaFuncSetAttribute(dynca_bwd_bf16_kernel<NS_, X3_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        dynca_bwd_bf16_kernel<NS_, X3_><<<grid, BB_THREADS, smem, s>>>(a);                                                        \
    } while (0)
    if (g.ns == 2) { if (x3) BB_LAUNCH(2, true); else BB_LAUNCH(2, false); }
    else { if (x3) BB_LAUNCH(1, true); else BB_LAUNCH(1, false); }
    NCA_LAUNCH_OK();
    return NCA_OK;
}

size_t dynca_bf16_coarse_floats(const DyncaGeom& g) {
    return g.ns == 2 ? nca_align_up((size_t)g.B * g.C * (g.H / 2) * (g.W / 2), 64) : 0;
}
