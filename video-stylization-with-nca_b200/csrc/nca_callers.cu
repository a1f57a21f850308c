// The callers either side of the NCA step: pool batch assembly and write-back, the fused
// gradient-normalise + Adam step, the overflow loss with its gradient, and the per-frame conditioning / rgb8 packing of the
// inference stream.  All of them are HBM-bound streaming kernels (or, for the optimizer, one latency-bound launch instead of
// ~40): 128-bit coalesced accesses, grids sized from the SM count, no tensor-core reshaping.
#include <math.h>
#include "nca_common.cuh"

int nca_check_device();   // nca_api.cu

namespace {

int sm_count() { return nca_sm_count(); }

// ------------------------------------------------------------------------------------------------------------------------
// Pool gather / scatter.  One (sample, channel) plane per blockIdx.y; blockIdx.x strides over the plane in float4.
// ------------------------------------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256) pool_gather_kernel(int N, int Cp, int Cx, size_t hw, const float* __restrict__ pool,
                                                          const int64_t* __restrict__ idx, const float* __restrict__ extra,
                                                          const float* __restrict__ seed_state, int inject_n,
                                                          const uint8_t* __restrict__ reseed, float* __restrict__ out) {
    const int C = Cp + Cx;
    const int b = blockIdx.y / C, c = blockIdx.y - b * C;
    const float* src = nullptr;          // nullptr = zeros
    if (c >= Cp) {
        src = extra + ((size_t)b * Cx + (c - Cp)) * hw;
    } else if (b < inject_n || (reseed && reseed[b])) {
        src = seed_state ? seed_state + (size_t)c * hw : nullptr;
    } else {
        const int64_t s = idx[b];
        src = (s >= 0 && s < N) ? pool + ((size_t)s * Cp + c) * hw : nullptr;     // out-of-range slot reads as zeros
    }
    float* dst = out + ((size_t)b * C + c) * hw;
    if (VEC) {
        const size_t n4 = hw >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
            d4[i] = src ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x)
            dst[i] = src ? __ldg(src + i) : 0.f;
    }
}

// flags[b] = 1 when pool[idx[b]] has no living cell: alive() = max_pool2d(x[:, d], 3, 1, 1) > thr (nca.py:152-163) is empty
// exactly when no x[:, d] exceeds thr
__global__ void __launch_bounds__(256) pool_dead_flags_kernel(int N, int Cp, size_t hw, const float* __restrict__ pool,
                                                              const int64_t* __restrict__ idx, int living_dim, float thr,
                                                              uint8_t* __restrict__ flags) {
    const int b = blockIdx.x;
    const int64_t s = idx[b];
    int alive = 0;
    if (s >= 0 && s < N) {
        const float* x = pool + ((size_t)s * Cp + living_dim) * hw;
        for (size_t i = threadIdx.x; i < hw; i += blockDim.x) alive |= __ldg(x + i) > thr;
    }
    alive = __syncthreads_or(alive);
    if (threadIdx.x == 0) flags[b] = alive ? 0 : 1;
}

template <bool VEC>
__global__ void __launch_bounds__(256) pool_scatter_kernel(int N, int Cp, int C, size_t hw, float* __restrict__ pool,
                                                           const int64_t* __restrict__ idx, const float* __restrict__ states) {
    const int b = blockIdx.y / Cp, c = blockIdx.y - b * Cp;
    const int64_t s = idx[b];
    if (s < 0 || s >= N) return;                                                   // out-of-range slot: nothing is written
    const float* src = states + ((size_t)b * C + c) * hw;
    float* dst = pool + ((size_t)s * Cp + c) * hw;
    if (VEC) {
        const size_t n4 = hw >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) d4[i] = __ldg(s4 + i);
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) dst[i] = __ldg(src + i);
    }
}

// blocks along a plane: enough that every SM has ~8 CTAs (2048 threads, ~100 KB of loads in flight) in total, at most one per
// 256 items (one float4 per thread: with few planes - a single 1080p frame - parallelism matters more than loop amortisation)
unsigned plane_blocks(size_t n_items, unsigned planes) {
    const size_t want = ((size_t)sm_count() * 8 + planes - 1) / planes;
    const size_t most = (n_items + 255) / 256;
    size_t g = want < most ? want : most;
    return (unsigned)(g < 1 ? 1 : g);
}

// ------------------------------------------------------------------------------------------------------------------------
// Per-parameter gradient normalisation + Adam, one CTA per parameter tensor.
// ------------------------------------------------------------------------------------------------------------------------
struct AdamSegs {
    float* p[NCA_ADAM_MAX_TENSORS];
    float* g[NCA_ADAM_MAX_TENSORS];
    float* m[NCA_ADAM_MAX_TENSORS];
    float* v[NCA_ADAM_MAX_TENSORS];
    long long n[NCA_ADAM_MAX_TENSORS];
};

__device__ __forceinline__ float block_sum(float v, float* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    float t = l < (int)(blockDim.x >> 5) ? sh[l] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    __syncthreads();
    return t;        // every thread holds the total
}

__global__ void __launch_bounds__(1024) normalized_adam_kernel(AdamSegs segs, float norm_eps, float one_minus_b1, float b2,
                                                               float one_minus_b2, float step_size, float bc2_sqrt, float eps,
                                                               int normalise, int zero_grads) {
    __shared__ float sh[32];
    const int s = blockIdx.x;
    float* __restrict__ p = segs.p[s];
    float* __restrict__ g = segs.g[s];
    float* __restrict__ m = segs.m[s];
    float* __restrict__ v = segs.v[s];
    const long long n = segs.n[s];
    float denom_g = 1.f;
    if (normalise) {
        float ss = 0.f;
        for (long long i = threadIdx.x; i < n; i += blockDim.x) { const float x = g[i]; ss = fmaf(x, x, ss); }
        denom_g = sqrtf(block_sum(ss, sh)) + norm_eps;                   // p.grad.norm() + eps   (experiments.py:253)
    }
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float gi = normalise ? __fdiv_rn(g[i], denom_g) : g[i];    // p.grad /= ...
        // torch.optim.Adam, single-tensor path (amsgrad off, weight_decay 0, maximize off)
        const float mi = m[i] + (gi - m[i]) * one_minus_b1;              // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = __fmul_rn(v[i], b2) + one_minus_b2 * gi * gi;   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float den = __fdiv_rn(sqrtf(vi), bc2_sqrt) + eps;          // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
        p[i] = p[i] - step_size * __fdiv_rn(mi, den);                    // param.addcdiv_(exp_avg, denom, value=-step_size)
        m[i] = mi;
        v[i] = vi;
        g[i] = zero_grads ? 0.f : gi;
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// Overflow loss: mean |x - clamp(x, -1, 1)| and its gradient, one pass over the state.
// ------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float overflow_term(float x, float gs, float* g) {
    const float over = fabsf(x - fminf(fmaxf(x, -1.f), 1.f));
    *g = x > 1.f ? gs : (x < -1.f ? -gs : 0.f);
    return over;
}

__global__ void __launch_bounds__(256) overflow_partial_kernel(const float* __restrict__ x, size_t n, float* __restrict__ partial,
                                                               float* __restrict__ grad, float gs, const float* __restrict__ gs_dev,
                                                               int accumulate) {
    __shared__ float sh[32];
    float acc = 0.f;
    if (gs_dev) gs *= __ldg(gs_dev);          // the incoming gradient of the loss, read on the device (no host round trip)
    const size_t n4 = n >> 2;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(grad)) & 15u) == 0;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
    if (vec) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        float4* g4 = reinterpret_cast<float4*>(grad);
        for (size_t i = tid; i < n4; i += nth) {
            const float4 a = __ldg(x4 + i);
            float4 d;
            acc += overflow_term(a.x, gs, &d.x);
            acc += overflow_term(a.y, gs, &d.y);
            acc += overflow_term(a.z, gs, &d.z);
            acc += overflow_term(a.w, gs, &d.w);
            if (grad) {
                if (accumulate) { const float4 o = g4[i]; d.x += o.x; d.y += o.y; d.z += o.z; d.w += o.w; }
                g4[i] = d;
            }
        }
        for (size_t i = (n4 << 2) + tid; i < n; i += nth) {
            float d;
            acc += overflow_term(x[i], gs, &d);
            if (grad) grad[i] = accumulate ? grad[i] + d : d;
        }
    } else {
        for (size_t i = tid; i < n; i += nth) {
            float d;
            acc += overflow_term(x[i], gs, &d);
            if (grad) grad[i] = accumulate ? grad[i] + d : d;
        }
    }
    const float t = block_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// fixed-order final reduction (bit-reproducible run to run): one CTA sums the per-CTA partials
__global__ void __launch_bounds__(1024) overflow_final_kernel(const float* __restrict__ partial, int n_partial, float inv_n,
                                                              float* __restrict__ loss) {
    __shared__ float sh[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n_partial; i += blockDim.x) acc += partial[i];
    const float t = block_sum(acc, sh);
    if (threadIdx.x == 0) loss[0] = t * inv_n;
}

// ------------------------------------------------------------------------------------------------------------------------
// Inference stream: conditioning channel from the frame, rgb8 packing of the state.
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) frame_gray_kernel(int C, int ch, size_t hw, const float* __restrict__ rgb,
                                                         float* __restrict__ state) {
    const int b = blockIdx.y;
    const float* r = rgb + (size_t)b * 3 * hw;
    float* dst = state + ((size_t)b * C + ch) * hw;
    // torch.mean(rgb, dim=1): (r + g + b) / 3 in fp32, summed in channel order (preprocess_texture.py:178-179)
    if ((hw & 3) == 0) {
        const size_t n4 = hw >> 2;
        const float4* r4 = reinterpret_cast<const float4*>(r);
        const float4* g4 = reinterpret_cast<const float4*>(r + hw);
        const float4* b4 = reinterpret_cast<const float4*>(r + 2 * hw);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            const float4 a = __ldg(r4 + i), bb = __ldg(g4 + i), c = __ldg(b4 + i);
            float4 o;
            o.x = __fdiv_rn(__fadd_rn(__fadd_rn(a.x, bb.x), c.x), 3.f);
            o.y = __fdiv_rn(__fadd_rn(__fadd_rn(a.y, bb.y), c.y), 3.f);
            o.z = __fdiv_rn(__fadd_rn(__fadd_rn(a.z, bb.z), c.z), 3.f);
            o.w = __fdiv_rn(__fadd_rn(__fadd_rn(a.w, bb.w), c.w), 3.f);
            d4[i] = o;
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x)
            dst[i] = __fdiv_rn(__fadd_rn(__fadd_rn(__ldg(r + i), __ldg(r + hw + i)), __ldg(r + 2 * hw + i)), 3.f);
    }
}

__device__ __forceinline__ uint32_t rgb8_of(float x, float scale) {
    float v = __fmul_rn(x, scale);                    // to_rgb: x[:, :3] * 2           (dynca.py:130-131)
    v = fminf(fmaxf(v, -1.f), 1.f);                   // np.clip(img, -1, 1)            (video_utils.py:80)
    v = __fmul_rn(__fadd_rn(v, 1.f), 0.5f);           // (img + 1) / 2                  (video_utils.py:81)
    v = __fmul_rn(fminf(fmaxf(v, 0.f), 1.f), 255.f);  // np.uint8(img.clip(0, 1) * 255) (video_utils.py:26): truncation
    return (uint32_t)v;
}

// four horizontally adjacent pixels per thread: three float4 plane loads, 12 bytes = three 32-bit stores (HWC, interleaved)
__global__ void __launch_bounds__(256) state_rgb8_kernel(int C, size_t hw, const float* __restrict__ state, float scale,
                                                         uint8_t* __restrict__ out) {
    const int b = blockIdx.y;
    const float* s = state + (size_t)b * C * hw;
    uint8_t* o = out + (size_t)b * 3 * hw;
    if ((hw & 3) == 0) {
        const size_t n4 = hw >> 2;
        const float4* r4 = reinterpret_cast<const float4*>(s);
        const float4* g4 = reinterpret_cast<const float4*>(s + hw);
        const float4* b4 = reinterpret_cast<const float4*>(s + 2 * hw);
        uint32_t* o32 = reinterpret_cast<uint32_t*>(o);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            const float4 r = __ldg(r4 + i), g = __ldg(g4 + i), bl = __ldg(b4 + i);
            const uint32_t r0 = rgb8_of(r.x, scale), g0 = rgb8_of(g.x, scale), b0 = rgb8_of(bl.x, scale);
            const uint32_t r1 = rgb8_of(r.y, scale), g1 = rgb8_of(g.y, scale), b1 = rgb8_of(bl.y, scale);
            const uint32_t r2 = rgb8_of(r.z, scale), g2 = rgb8_of(g.z, scale), b2 = rgb8_of(bl.z, scale);
            const uint32_t r3 = rgb8_of(r.w, scale), g3 = rgb8_of(g.w, scale), b3 = rgb8_of(bl.w, scale);
            o32[3 * i + 0] = r0 | (g0 << 8) | (b0 << 16) | (r1 << 24);
            o32[3 * i + 1] = g1 | (b1 << 8) | (r2 << 16) | (g2 << 24);
            o32[3 * i + 2] = b2 | (r3 << 8) | (g3 << 16) | (b3 << 24);
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (size_t)gridDim.x * blockDim.x) {
            o[3 * i + 0] = (uint8_t)rgb8_of(__ldg(s + i), scale);
            o[3 * i + 1] = (uint8_t)rgb8_of(__ldg(s + hw + i), scale);
            o[3 * i + 2] = (uint8_t)rgb8_of(__ldg(s + 2 * hw + i), scale);
        }
    }
}

}  // namespace

#define NCA_ALIGNED16(p) ((((uintptr_t)(p)) & 15u) == 0)

extern "C" {

int nca_pool_gather(int32_t N, int32_t Cp, int32_t H, int32_t W, const float* pool, const int64_t* idx, int32_t B,
                    const float* extra, int32_t Cx, const float* seed_state, int32_t inject_n, const uint8_t* reseed_flags,
                    float* out, void* stream) {
    NCA_CHECK_ARG(N > 0 && Cp > 0 && H > 0 && W > 0 && B > 0 && Cx >= 0, "nca_pool_gather: bad sizes N=%d Cp=%d H=%d W=%d B=%d Cx=%d", N, Cp, H, W, B, Cx);
    NCA_CHECK_ARG(pool && idx && out, "nca_pool_gather: pool, idx and out must not be NULL");
    NCA_CHECK_ARG(Cx == 0 || extra, "nca_pool_gather: Cx=%d extra channels need an extra pointer", Cx);
    NCA_CHECK_ARG(inject_n >= 0 && inject_n <= B, "nca_pool_gather: inject_n=%d outside 0..B", inject_n);
    NCA_CHECK_ARG((long long)B * (Cp + Cx) <= 65535, "nca_pool_gather: B*(Cp+Cx)=%lld planes exceed 65535", (long long)B * (Cp + Cx));
    int rc = nca_check_device();
    if (rc) return rc;
    const size_t hw = (size_t)H * W;
    const bool vec = (hw & 3) == 0 && NCA_ALIGNED16(pool) && NCA_ALIGNED16(out) && (!extra || NCA_ALIGNED16(extra)) && (!seed_state || NCA_ALIGNED16(seed_state));
    const unsigned planes = (unsigned)(B * (Cp + Cx));
    dim3 grid(plane_blocks(vec ? hw >> 2 : hw, planes), planes);
    cudaStream_t s = (cudaStream_t)stream;
    if (vec) pool_gather_kernel<true><<<grid, 256, 0, s>>>(N, Cp, Cx, hw, pool, idx, extra, seed_state, inject_n, reseed_flags, out);
    else pool_gather_kernel<false><<<grid, 256, 0, s>>>(N, Cp, Cx, hw, pool, idx, extra, seed_state, inject_n, reseed_flags, out);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int nca_pool_dead_flags(int32_t N, int32_t Cp, int32_t H, int32_t W, const float* pool, const int64_t* idx, int32_t B,
                        int32_t living_dim, float alive_thr, uint8_t* flags, void* stream) {
    NCA_CHECK_ARG(N > 0 && Cp > 0 && H > 0 && W > 0 && B > 0, "nca_pool_dead_flags: bad sizes N=%d Cp=%d H=%d W=%d B=%d", N, Cp, H, W, B);
    NCA_CHECK_ARG(living_dim >= 0 && living_dim < Cp, "nca_pool_dead_flags: living_dim=%d outside 0..%d", living_dim, Cp - 1);
    NCA_CHECK_ARG(pool && idx && flags, "nca_pool_dead_flags: NULL pointer");
    int rc = nca_check_device();
    if (rc) return rc;
    pool_dead_flags_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(N, Cp, (size_t)H * W, pool, idx, living_dim, alive_thr, flags);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int nca_pool_scatter(int32_t N, int32_t Cp, int32_t H, int32_t W, float* pool, const int64_t* idx, int32_t B,
                     const float* states, int32_t C, void* stream) {
    NCA_CHECK_ARG(N > 0 && Cp > 0 && H > 0 && W > 0 && B > 0 && C >= Cp, "nca_pool_scatter: bad sizes N=%d Cp=%d H=%d W=%d B=%d C=%d", N, Cp, H, W, B, C);
    NCA_CHECK_ARG(pool && idx && states, "nca_pool_scatter: pool, idx and states must not be NULL");
    NCA_CHECK_ARG((long long)B * Cp <= 65535, "nca_pool_scatter: B*Cp=%lld planes exceed 65535", (long long)B * Cp);
    int rc = nca_check_device();
    if (rc) return rc;
    const size_t hw = (size_t)H * W;
    const bool vec = (hw & 3) == 0 && NCA_ALIGNED16(pool) && NCA_ALIGNED16(states);
    const unsigned planes = (unsigned)(B * Cp);
    dim3 grid(plane_blocks(vec ? hw >> 2 : hw, planes), planes);
    cudaStream_t s = (cudaStream_t)stream;
    if (vec) pool_scatter_kernel<true><<<grid, 256, 0, s>>>(N, Cp, C, hw, pool, idx, states);
    else pool_scatter_kernel<false><<<grid, 256, 0, s>>>(N, Cp, C, hw, pool, idx, states);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int nca_normalized_adam_step(int32_t n_tensors, float* const* params, float* const* grads, float* const* exp_avg,
                             float* const* exp_avg_sq, const int64_t* numel, int32_t step, float lr, float beta1, float beta2,
                             float eps, float norm_eps, int32_t normalise, int32_t zero_grads, void* stream) {
    NCA_CHECK_ARG(n_tensors > 0 && n_tensors <= NCA_ADAM_MAX_TENSORS, "nca_normalized_adam_step: n_tensors=%d outside 1..%d", n_tensors, NCA_ADAM_MAX_TENSORS);
    NCA_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && numel, "nca_normalized_adam_step: NULL table");
    NCA_CHECK_ARG(step >= 1, "nca_normalized_adam_step: step=%d must be >= 1 (torch.optim.Adam counts from 1)", step);
    NCA_CHECK_ARG(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "nca_normalized_adam_step: bad hyper-parameters");
    AdamSegs segs;
    memset(&segs, 0, sizeof(segs));
    for (int i = 0; i < n_tensors; ++i) {
        NCA_CHECK_ARG(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i] && numel[i] > 0, "nca_normalized_adam_step: tensor %d has a NULL pointer or no elements", i);
        segs.p[i] = params[i]; segs.g[i] = grads[i]; segs.m[i] = exp_avg[i]; segs.v[i] = exp_avg_sq[i]; segs.n[i] = numel[i];
    }
    int rc = nca_check_device();
    if (rc) return rc;
    // torch/optim/adam.py _single_tensor_adam: python-float bias corrections
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float bc2_sqrt = (float)sqrt(bc2);
    normalized_adam_kernel<<<n_tensors, 1024, 0, (cudaStream_t)stream>>>(segs, norm_eps, (float)(1.0 - (double)beta1), beta2,
                                                                        (float)(1.0 - (double)beta2), step_size, bc2_sqrt, eps,
                                                                        normalise, zero_grads);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

size_t nca_overflow_workspace_bytes(void) { return 1024 * sizeof(float); }

int nca_overflow_loss(const float* x, size_t n, float* loss_out, float* grad_out, float grad_scale, const float* grad_scale_dev,
                      int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream) {
    NCA_CHECK_ARG(x && loss_out && n > 0, "nca_overflow_loss: x / loss_out NULL or n == 0");
    if (!workspace || workspace_bytes < nca_overflow_workspace_bytes()) {
        nca_set_error("nca_overflow_loss: workspace of %zu bytes needed, got %zu", nca_overflow_workspace_bytes(), workspace_bytes);
        return NCA_ERR_WORKSPACE;
    }
    int rc = nca_check_device();
    if (rc) return rc;
    size_t blocks = (n / 4 + 255) / 256;
    const size_t cap = (size_t)sm_count() * 6 < 1024 ? (size_t)sm_count() * 6 : 1024;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    float* partial = (float*)workspace;
    cudaStream_t s = (cudaStream_t)stream;
    // d/dx mean|x - clamp(x)| = sign(x) [|x| > 1] / n
    overflow_partial_kernel<<<(unsigned)blocks, 256, 0, s>>>(x, n, partial, grad_out, (float)((double)grad_scale / (double)n), grad_scale_dev, accumulate);
    NCA_LAUNCH_OK();
    overflow_final_kernel<<<1, 1024, 0, s>>>(partial, (int)blocks, (float)(1.0 / (double)n), loss_out);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int nca_frame_to_cond_channel(int32_t B, int32_t C, int32_t H, int32_t W, const float* frame_rgb, float* state, int32_t ch, void* stream) {
    NCA_CHECK_ARG(B > 0 && B <= 65535 && C > 0 && H > 0 && W > 0, "nca_frame_to_cond_channel: bad sizes B=%d C=%d H=%d W=%d", B, C, H, W);
    NCA_CHECK_ARG(ch >= 0 && ch < C, "nca_frame_to_cond_channel: channel %d outside 0..%d", ch, C - 1);
    NCA_CHECK_ARG(frame_rgb && state, "nca_frame_to_cond_channel: NULL pointer");
    NCA_CHECK_ARG(NCA_ALIGNED16(frame_rgb) && NCA_ALIGNED16(state), "nca_frame_to_cond_channel: pointers must be 16-byte aligned");
    int rc = nca_check_device();
    if (rc) return rc;
    const size_t hw = (size_t)H * W;
    dim3 grid(plane_blocks((hw & 3) == 0 ? hw >> 2 : hw, (unsigned)B), (unsigned)B);
    frame_gray_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(C, ch, hw, frame_rgb, state);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

int nca_state_to_rgb8(int32_t B, int32_t C, int32_t H, int32_t W, const float* state, float scale, uint8_t* out_hwc, void* stream) {
    NCA_CHECK_ARG(B > 0 && B <= 65535 && C >= 3 && H > 0 && W > 0, "nca_state_to_rgb8: bad sizes B=%d C=%d H=%d W=%d (C >= 3)", B, C, H, W);
    NCA_CHECK_ARG(state && out_hwc, "nca_state_to_rgb8: NULL pointer");
    NCA_CHECK_ARG(NCA_ALIGNED16(state) && NCA_ALIGNED16(out_hwc), "nca_state_to_rgb8: pointers must be 16-byte aligned");
    int rc = nca_check_device();
    if (rc) return rc;
    const size_t hw = (size_t)H * W;
    dim3 grid(plane_blocks((hw & 3) == 0 ? hw >> 2 : hw, (unsigned)B), (unsigned)B);
    state_rgb8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(C, hw, state, scale, out_hwc);
    NCA_LAUNCH_OK();
    return NCA_OK;
}

}  // extern "C"
