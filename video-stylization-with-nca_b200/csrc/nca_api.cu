// C ABI of libnca_b200.so (include/nca_b200.h): argument checking, workspace carving, step loops.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "nca_internal.h"

static thread_local char t_err[512] = "";
static std::atomic<long long> g_launches{0};   // process-wide: autograd runs backward on its own thread

void nca_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}
void nca_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" {

const char* nca_last_error(void) { return t_err; }
int nca_abi_version(void) { return NCA_B200_ABI_VERSION; }
long long nca_launch_count(void) { return g_launches.load(); }
void nca_launch_count_reset(void) { g_launches.store(0); }

static int check_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        nca_set_error("no CUDA device available (%s); libnca_b200 has no CPU fallback", cudaGetErrorString(e));
        cudaGetLastError();
        return NCA_ERR_CUDA;
    }
    return NCA_OK;
}
#define NCA_ALIGNED16(p) ((((uintptr_t)(p)) & 15u) == 0)

static FireMask make_mask(const NcaDyncaDesc* d, const DyncaGeom& g, const float* masks, uint64_t seed, int t) {
    FireMask m;
    m.supplied = d->mask_mode == NCA_MASK_SUPPLIED ? masks + (size_t)t * g.B * g.H * g.W : nullptr;
    m.k0 = (uint32_t)(seed & 0xffffffffu);
    m.k1 = (uint32_t)(seed >> 32);
    m.t = 0;
    m.thr = nca_fire_threshold(d->update_rate, 0);
    return m;
}

static int check_common(const NcaDyncaDesc* d, DyncaGeom* g, const float* cond, const float* masks) {
    int rc = dynca_make_geom(d, g);
    if (rc) return rc;
    NCA_CHECK_ARG(d->precision == NCA_PREC_FP32 || d->precision == NCA_PREC_BF16, "bad precision %d", d->precision);
    NCA_CHECK_ARG(g->cond_kind != NCA_COND_TENSOR || cond != nullptr, "cond_kind == TENSOR needs a cond pointer");
    NCA_CHECK_ARG(d->mask_mode != NCA_MASK_SUPPLIED || masks != nullptr, "mask_mode == SUPPLIED needs a masks pointer");
    return check_device();
}

size_t nca_dynca_workspace_bytes(const NcaDyncaDesc* d, int32_t backward) {
    DyncaGeom g;
    if (dynca_make_geom(d, &g)) return 0;
    // layout: [fp32 padded weights | (backward) grad accumulators, 2 state-gradient buffers | (bf16) operand images]
    size_t n = dynca_f32_weight_floats(g);
    if (backward) n += dynca_f32_grad_floats(g) + 2 * nca_align_up((size_t)g.B * g.C * g.H * g.W, 64);
    size_t bytes = n * sizeof(float);
    if (d->precision == NCA_PREC_BF16) {
        size_t b = backward ? dynca_bf16_bwd_weight_bytes(g) : dynca_bf16_weight_bytes(g);
        if (b == 0) return 0;
        bytes += b + dynca_bf16_coarse_floats(g) * sizeof(float);   // + coarse (2x2-mean) state of the current step
    }
    return bytes;
}

int nca_dynca_perceive(const NcaDyncaDesc* d, const float* x, const float* cond, float* z, void* stream) {
    DyncaGeom g;
    NcaDyncaDesc dd;
    NCA_CHECK_ARG(d != nullptr, "desc is NULL");
    dd = *d; dd.mask_mode = NCA_MASK_PHILOX;
    int rc = check_common(&dd, &g, cond, nullptr);
    if (rc) return rc;
    NCA_CHECK_ARG(x && z, "x / z is NULL");
    return dynca_f32_perceive(g, x, cond, z, (cudaStream_t)stream);
}

int nca_edge_extract(int B, int H, int W, const float* img, int tanh_transform, float* out, void* stream) {
    NCA_CHECK_ARG(B > 0 && H > 0 && W > 0 && img && out, "bad edge_extract arguments");
    int rc = check_device();
    if (rc) return rc;
    return nca_edge_extract_launch(B, H, W, img, tanh_transform, out, (cudaStream_t)stream);
}

int nca_philox_mask(int32_t B, int32_t H, int32_t W, float rate, int32_t enc, uint64_t seed, int32_t t0, int32_t T,
                    float* out, void* stream) {
    NCA_CHECK_ARG(B > 0 && H > 0 && W > 0 && T > 0 && out, "bad philox_mask arguments");
    int rc = check_device();
    if (rc) return rc;
    return nca_philox_mask_launch(B, H, W, rate, enc, seed, t0, T, out, (cudaStream_t)stream);
}

int nca_dynca_forward(const NcaDyncaDesc* d, const NcaDyncaWeights* w, const float* cond, const float* masks,
                      uint64_t seed, int32_t t0, int32_t T, int32_t keep_history, float* states, void* workspace,
                      size_t workspace_bytes, void* stream) {
    DyncaGeom g;
    int rc = check_common(d, &g, cond, masks);
    if (rc) return rc;
    NCA_CHECK_ARG(w && w->w1 && w->b1 && w->w2 && w->b2, "weights are NULL");
    NCA_CHECK_ARG(states != nullptr && T >= 0, "states is NULL or T < 0");
    NCA_CHECK_ARG(NCA_ALIGNED16(states) && NCA_ALIGNED16(workspace), "states / workspace must be 16-byte aligned");
    if (workspace == nullptr || workspace_bytes < nca_dynca_workspace_bytes(d, 0)) {
        nca_set_error("workspace too small: %zu < %zu", workspace_bytes, nca_dynca_workspace_bytes(d, 0));
        return NCA_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    float* wsW = (float*)workspace;
    void* wsB = (uint8_t*)workspace + dynca_f32_weight_floats(g) * sizeof(float);
    const bool bf16 = d->precision == NCA_PREC_BF16;
    float* wsXc = bf16 ? (float*)((uint8_t*)wsB + dynca_bf16_weight_bytes(g)) : nullptr;
    rc = bf16 ? dynca_bf16_prep_weights(g, w, wsB, s) : dynca_f32_prep_weights(g, w, wsW, s);
    if (rc) return rc;
    const size_t n = (size_t)g.B * g.C * g.H * g.W;
    for (int t = 0; t < T; ++t) {
        FireMask fm = make_mask(d, g, masks, seed, t);
        fm.t = (uint32_t)(t0 + t);
        const float* xin = keep_history ? states + (size_t)t * n : states + (size_t)(t & 1) * n;
        float* xout = keep_history ? states + (size_t)(t + 1) * n : states + (size_t)((t + 1) & 1) * n;
        rc = bf16 ? dynca_bf16_forward_step(g, wsB, wsXc, xin, xout, cond, fm, s) : dynca_f32_forward_step(g, wsW, xin, xout, cond, fm, s);
        if (rc) return rc;
    }
    return NCA_OK;
}

int nca_dynca_backward(const NcaDyncaDesc* d, const NcaDyncaWeights* w, const float* cond, const float* masks,
                       uint64_t seed, int32_t t0, int32_t T, const float* states, const float* g_final,
                       const float* const* g_taps, const int32_t* tap_steps, int32_t n_taps, int32_t tap_c,
                       float tap_scale, float* gx0, const NcaDyncaWeightGrads* gw, void* workspace,
                       size_t workspace_bytes, void* stream) {
    DyncaGeom g;
    int rc = check_common(d, &g, cond, masks);
    if (rc) return rc;
    NCA_CHECK_ARG(w && w->w1 && w->b1 && w->w2 && w->b2, "weights are NULL");
    NCA_CHECK_ARG(gw && gw->w1 && gw->b1 && gw->w2 && gw->b2, "weight-gradient outputs are NULL");
    NCA_CHECK_ARG(states != nullptr && gx0 != nullptr && T >= 0, "states / gx0 is NULL or T < 0");
    NCA_CHECK_ARG(n_taps >= 0 && (n_taps == 0 || (g_taps && tap_steps && tap_c > 0 && tap_c <= g.C)), "bad tap arguments (n_taps=%d, tap_c=%d)", n_taps, tap_c);
    for (int i = 0; i < n_taps; ++i)
        NCA_CHECK_ARG(g_taps[i] && tap_steps[i] >= 1 && tap_steps[i] <= T && (i == 0 || tap_steps[i] > tap_steps[i - 1]),
                      "tap_steps must be strictly increasing in 1..T with non-NULL gradients (entry %d)", i);
    NCA_CHECK_ARG(NCA_ALIGNED16(states) && NCA_ALIGNED16(workspace), "states / workspace must be 16-byte aligned");
    if (workspace == nullptr || workspace_bytes < nca_dynca_workspace_bytes(d, 1)) {
        nca_set_error("workspace too small: %zu < %zu", workspace_bytes, nca_dynca_workspace_bytes(d, 1));
        return NCA_ERR_WORKSPACE;
    }
    // NCA_PREC_BF16: tcgen05 BPTT kernel when the shape is supported (fc % 32 == 0, fc <= 128), else the fp32 kernel
    const bool bf16 = d->precision == NCA_PREC_BF16 && dynca_bf16_bwd_supported(g);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)g.B * g.C * g.H * g.W, nb = n * sizeof(float);
    float* wsW = (float*)workspace;
    float* wsG = wsW + dynca_f32_weight_floats(g);
    float* gbuf[2] = {wsG + dynca_f32_grad_floats(g), wsG + dynca_f32_grad_floats(g) + nca_align_up(n, 64)};
    void* wsB = (void*)(gbuf[1] + nca_align_up(n, 64));
    float* wsXc = (float*)((uint8_t*)wsB + dynca_bf16_bwd_weight_bytes(g));
    rc = bf16 ? dynca_bf16_prep_bwd_weights(g, w, wsB, s) : dynca_f32_prep_weights(g, w, wsW, s);
    if (rc) return rc;
    NCA_CUDA_OK(cudaMemsetAsync(wsG, 0, dynca_f32_grad_floats(g) * sizeof(float), s));
    int ti = n_taps - 1;   // taps are consumed from the last step backwards
    if (T == 0) {
        // gx0 = g_final + tap at states[0] is not defined (taps are for t = 1..T): plain copy
        if (g_final) NCA_CUDA_OK(cudaMemcpyAsync(gx0, g_final, nb, cudaMemcpyDeviceToDevice, s));
        else NCA_CUDA_OK(cudaMemsetAsync(gx0, 0, nb, s));
    }
    const float* gnext = g_final;   // NULL = zeros
    for (int t = T - 1; t >= 0; --t) {
        FireMask fm = make_mask(d, g, masks, seed, t);
        fm.t = (uint32_t)(t0 + t);
        float* gout = t == 0 ? gx0 : gbuf[t & 1];
        NCA_CUDA_OK(cudaMemsetAsync(gout, 0, nb, s));
        const float* tap = nullptr;   // gradient injected at states[t+1]
        if (ti >= 0 && tap_steps[ti] == t + 1) tap = g_taps[ti--];
        rc = bf16 ? dynca_bf16_backward_step(g, wsB, wsXc, wsG, states + (size_t)t * n, gnext, tap, tap_c, tap_scale, gout, cond, fm, s)
                  : dynca_f32_backward_step(g, wsW, wsG, states + (size_t)t * n, gnext, tap, tap_c, tap_scale, gout, cond, fm, s);
        if (rc) return rc;
        gnext = gout;
    }
    return dynca_f32_unpack_grads(g, wsG, gw, s);
}

/* ---- ConditionedNCA: implemented in enc_f32.cu ---- */

}  // extern "C"
