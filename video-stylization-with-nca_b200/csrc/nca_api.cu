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
int nca_device_ordinal() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev;
}
int nca_sm_count() {
    static std::atomic<int> cache[NCA_MAX_DEVICES];      // zero-initialised; racing writers store the same value
    const int dev = nca_device_ordinal();
    int n = dev < NCA_MAX_DEVICES ? cache[dev].load(std::memory_order_relaxed) : 0;
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        if (dev < NCA_MAX_DEVICES) cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

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
}  // extern "C"
int nca_check_device() { return check_device(); }   // for the other translation units (nca_callers.cu)
extern "C" {

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
    NCA_CHECK_ARG(d->precision == NCA_PREC_FP32 || d->precision == NCA_PREC_BF16 || d->precision == NCA_PREC_F16X3, "bad precision %d", d->precision);
    NCA_CHECK_ARG(g->cond_kind != NCA_COND_TENSOR || cond != nullptr, "cond_kind == TENSOR needs a cond pointer");
    NCA_CHECK_ARG(d->mask_mode != NCA_MASK_SUPPLIED || masks != nullptr, "mask_mode == SUPPLIED needs a masks pointer");
    return check_device();
}

// kernel variant of a description: 0 fp32 CUDA cores, 1 tcgen05 (4x32 tiles, cp.async), 2 tcgen05 (8x16 tiles, TMA),
// 3 tcgen05 with split-precision operands (NCA_PREC_F16X3: the 4x32-tile kernels with hi + lo images)
static int dynca_variant(const NcaDyncaDesc* d, const DyncaGeom& g, int backward) {
    if (d->precision == NCA_PREC_F16X3) {
        if (backward) return dynca_bf16_bwd_supported(g, true) ? 3 : 0;
        return dynca_bf16_supported(g, true) ? 3 : 0;
    }
    if (d->precision != NCA_PREC_BF16) return 0;
    if (backward) return dynca_tc2_bwd_supported(g) ? 2 : (dynca_bf16_bwd_supported(g) ? 1 : 0);
    return dynca_tc2_supported(g) ? 2 : 1;
}
int nca_dynca_kernel_variant(const NcaDyncaDesc* d, int32_t backward) {
    DyncaGeom g;
    if (dynca_make_geom(d, &g)) return -1;
    if (d->precision != NCA_PREC_FP32 && d->precision != NCA_PREC_BF16 && d->precision != NCA_PREC_F16X3) return -1;
    return dynca_variant(d, g, backward);
}

size_t nca_dynca_workspace_bytes(const NcaDyncaDesc* d, int32_t backward) {
    DyncaGeom g;
    if (dynca_make_geom(d, &g)) return 0;
    // layout: [fp32 padded weights | (backward) grad accumulators, 2 state-gradient buffers | (bf16) operand images |
    //          2 coarse-state slots]
    size_t n = dynca_f32_weight_floats(g);
    if (backward) n += dynca_f32_grad_floats(g) + 2 * nca_align_up((size_t)g.B * g.C * g.H * g.W, 64);
    size_t bytes = n * sizeof(float);
    if (d->precision == NCA_PREC_BF16 || d->precision == NCA_PREC_F16X3) {
        const bool x3 = d->precision == NCA_PREC_F16X3;
        size_t b = backward ? dynca_bf16_bwd_weight_bytes(g, x3) : dynca_bf16_weight_bytes(g, x3);
        if (b == 0) return 0;
        const size_t b2 = x3 ? 0 : (backward ? dynca_tc2_bwd_weight_bytes(g) : dynca_tc2_weight_bytes(g));
        // coarse slots: forward 2 (ping-pong states); backward 3 (coarse state of the step + 2 coarse-gradient buffers)
        bytes += (b > b2 ? b : b2) + (backward ? 3 : 2) * dynca_bf16_coarse_floats(g) * sizeof(float);
        // one step of perception operands for the tcgen05 BPTT when the caller kept no operand history (recompute path)
        if (backward && !x3 && dynca_tc2_bwd_supported(g)) bytes = nca_align_up(bytes, 256) + dynca_tc2_op_hist_bytes(g, 1);
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
    return nca_philox_mask_launch(B, H, W, rate, enc, seed, t0, nullptr, T, out, (cudaStream_t)stream);
}

int nca_philox_mask_at(int32_t B, int32_t H, int32_t W, float rate, int32_t enc, uint64_t seed, int32_t t0, const uint32_t* t0_dev,
                       int32_t T, float* out, void* stream) {
    NCA_CHECK_ARG(B > 0 && H > 0 && W > 0 && T > 0 && out, "bad philox_mask arguments");
    int rc = check_device();
    if (rc) return rc;
    return nca_philox_mask_launch(B, H, W, rate, enc, seed, t0, t0_dev, T, out, (cudaStream_t)stream);
}

size_t nca_dynca_op_hist_bytes(const NcaDyncaDesc* d, int32_t T) {
    DyncaGeom g;
    if (d == nullptr || T <= 0 || dynca_make_geom(d, &g)) return 0;
    if (d->precision != NCA_PREC_BF16) return 0;
    if (dynca_variant(d, g, 0) != 2 || dynca_variant(d, g, 1) != 2) return 0;
    return dynca_tc2_op_hist_bytes(g, T);
}

int nca_dynca_forward(const NcaDyncaDesc* d, const NcaDyncaWeights* w, const float* cond, const float* masks,
                      uint64_t seed, int32_t t0, int32_t T, int32_t keep_history, float* states, float* coarse_hist,
                      void* op_hist, void* workspace, size_t workspace_bytes, void* stream) {
    DyncaGeom g;
    int rc = check_common(d, &g, cond, masks);
    if (rc) return rc;
    NCA_CHECK_ARG(w && w->w1 && w->b1 && w->w2 && w->b2, "weights are NULL");
    NCA_CHECK_ARG(states != nullptr && T >= 0, "states is NULL or T < 0");
    NCA_CHECK_ARG(NCA_ALIGNED16(states) && NCA_ALIGNED16(workspace) && NCA_ALIGNED16(coarse_hist), "states / coarse_hist / workspace must be 16-byte aligned");
    {
        const size_t need = nca_dynca_workspace_bytes(d, 0);
        if (need == 0) {      // 0 = "this description cannot run" (e.g. a tensor-core precision with fc % 16 != 0 or more than 6 cond channels)
            nca_set_error("unsupported geometry for precision %d (fc %d, cond channels %d)", d->precision, g.fc, g.cc);
            return NCA_ERR_UNSUPPORTED;
        }
        if (workspace == nullptr || workspace_bytes < need) {
            nca_set_error("workspace too small: %zu < %zu", workspace_bytes, need);
            return NCA_ERR_WORKSPACE;
        }
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int variant = dynca_variant(d, g, 0);
    const bool x3 = d->precision == NCA_PREC_F16X3;
    const bool gen1 = variant == 1 || variant == 3;       // the 4x32-tile kernels (variant 3: with hi + lo operand images)
    float* wsW = (float*)workspace;
    void* wsB = (uint8_t*)workspace + dynca_f32_weight_floats(g) * sizeof(float);
    const size_t imgb = x3 ? dynca_bf16_weight_bytes(g, true)
                           : (dynca_bf16_weight_bytes(g) > dynca_tc2_weight_bytes(g) ? dynca_bf16_weight_bytes(g) : dynca_tc2_weight_bytes(g));
    float* wsXc = variant ? (float*)((uint8_t*)wsB + imgb) : nullptr;      // 2 coarse slots
    const size_t n = (size_t)g.B * g.C * g.H * g.W, nc = dynca_bf16_coarse_floats(g);
    if (T == 0) return NCA_OK;
    rc = variant == 2 ? dynca_tc2_prep_weights(g, w, wsB, s) : (gen1 ? dynca_bf16_prep_weights(g, w, wsB, s, x3) : dynca_f32_prep_weights(g, w, wsW, s));
    if (rc) return rc;
    // coarse states (two perception scales on the tensor-core paths): a history when the caller keeps one, else ping-pong
    const bool chist = keep_history && coarse_hist != nullptr && g.ns == 2;
    const bool ohist = keep_history && op_hist != nullptr && variant == 2 && dynca_tc2_bwd_supported(g);
    NCA_CHECK_ARG(NCA_ALIGNED16(op_hist), "op_hist must be 16-byte aligned");
    const size_t op_step = ohist ? dynca_tc2_op_hist_bytes(g, 1) : 0;
    float* cbase = chist ? coarse_hist : wsXc;
    const size_t cstride = chist ? (size_t)g.B * g.C * (g.H / 2) * (g.W / 2) : nc;
    DyncaTc2Maps maps;
    if (variant == 2) {
        NCA_CHECK_ARG(NCA_ALIGNED16(cond), "cond must be 16-byte aligned");
        rc = dynca_tc2_make_maps(g, states, keep_history ? T + 1 : 2, cbase, chist ? T + 1 : 2, cstride, cond, &maps);
        if (rc) return rc;
    }
    for (int t = 0; t < T; ++t) {
        FireMask fm = make_mask(d, g, masks, seed, t);
        fm.t = (uint32_t)(t0 + t);
        const int si = keep_history ? t : (t & 1), so = keep_history ? t + 1 : ((t + 1) & 1);
        const int ci = chist ? t : (t & 1), co = chist ? t + 1 : ((t + 1) & 1);
        const float* xin = states + (size_t)si * n;
        float* xout = states + (size_t)so * n;
        if (variant == 2) {
            if (g.ns == 2 && t == 0) { rc = dynca_bf16_coarsen(g, xin, cbase + (size_t)ci * cstride, s); if (rc) return rc; }
            const bool need_next = g.ns == 2 && (t + 1 < T || chist);
            rc = dynca_tc2_forward_step(g, wsB, &maps, si, xin, xout, ci, g.ns == 2 ? cbase + (size_t)ci * cstride : nullptr,
                                        need_next ? cbase + (size_t)co * cstride : nullptr, cond, fm, s, t > 0,
                                        ohist ? (uint8_t*)op_hist + (size_t)t * op_step : nullptr);
        } else if (gen1) {
            float* xc = g.ns == 2 ? cbase + (size_t)ci * cstride : nullptr;
            rc = dynca_bf16_forward_step(g, wsB, xc, xin, xout, cond, fm, s, x3);
            if (!rc && chist && t + 1 == T) rc = dynca_bf16_coarsen(g, xout, cbase + (size_t)co * cstride, s);
        } else {
            rc = dynca_f32_forward_step(g, wsW, xin, xout, cond, fm, s);
        }
        if (rc) return rc;
    }
    return NCA_OK;
}

int nca_dynca_backward(const NcaDyncaDesc* d, const NcaDyncaWeights* w, const float* cond, const float* masks,
                       uint64_t seed, int32_t t0, int32_t T, const float* states, const float* coarse_hist,
                       const void* op_hist, const float* g_final, const float* const* g_taps, const int32_t* tap_steps, int32_t n_taps,
                       int32_t tap_c, float tap_scale, float* gx0, const NcaDyncaWeightGrads* gw, void* workspace,
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
    {
        const size_t need = nca_dynca_workspace_bytes(d, 1);
        if (need == 0) {
            nca_set_error("unsupported geometry for precision %d (fc %d, cond channels %d)", d->precision, g.fc, g.cc);
            return NCA_ERR_UNSUPPORTED;
        }
        if (workspace == nullptr || workspace_bytes < need) {
            nca_set_error("workspace too small: %zu < %zu", workspace_bytes, need);
            return NCA_ERR_WORKSPACE;
        }
    }
    // NCA_PREC_BF16: tcgen05 BPTT kernels when the shape is supported (variant 2: 8x16 tiles + TMA, variant 1: 4x32 tiles),
    // else the fp32 kernel
    const int variant = dynca_variant(d, g, 1);
    const bool bf16 = variant != 0;
    const bool x3 = d->precision == NCA_PREC_F16X3;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)g.B * g.C * g.H * g.W, nb = n * sizeof(float);
    float* wsW = (float*)workspace;
    float* wsG = wsW + dynca_f32_weight_floats(g);
    float* gbuf[2] = {wsG + dynca_f32_grad_floats(g), wsG + dynca_f32_grad_floats(g) + nca_align_up(n, 64)};
    void* wsB = (void*)(gbuf[1] + nca_align_up(n, 64));
    const size_t imgb = x3 ? dynca_bf16_bwd_weight_bytes(g, true)
                           : (dynca_bf16_bwd_weight_bytes(g) > dynca_tc2_bwd_weight_bytes(g) ? dynca_bf16_bwd_weight_bytes(g) : dynca_tc2_bwd_weight_bytes(g));
    float* wsXc = (float*)((uint8_t*)wsB + imgb);
    const size_t nc = dynca_bf16_coarse_floats(g), ncx = (size_t)g.B * g.C * (g.H / 2) * (g.W / 2);
    rc = variant == 2 ? dynca_tc2_prep_bwd_weights(g, w, wsB, s) : (bf16 ? dynca_bf16_prep_bwd_weights(g, w, wsB, s, x3) : dynca_f32_prep_weights(g, w, wsW, s));
    if (rc) return rc;
    NCA_CUDA_OK(cudaMemsetAsync(wsG, 0, dynca_f32_grad_floats(g) * sizeof(float), s));
    int ti = n_taps - 1;   // taps are consumed from the last step backwards
    if (T == 0) {
        // gx0 = g_final + tap at states[0] is not defined (taps are for t = 1..T): plain copy
        if (g_final) NCA_CUDA_OK(cudaMemcpyAsync(gx0, g_final, nb, cudaMemcpyDeviceToDevice, s));
        else NCA_CUDA_OK(cudaMemsetAsync(gx0, 0, nb, s));
    }
    if (variant == 2 && T > 0) {
        // ping-pong gradient buffers F[p] (fine) / G[p] (coarse part); each launch zeroes the tile of its input it consumed,
        // so only the initial state needs memsets
        float* F[2] = {gbuf[0], gbuf[1]};
        float* G[2] = {wsXc + nc, wsXc + 2 * nc};
        NCA_CUDA_OK(cudaMemsetAsync(F[0], 0, nb, s));
        NCA_CUDA_OK(cudaMemsetAsync(F[1], 0, nb, s));
        NCA_CUDA_OK(cudaMemsetAsync(gx0, 0, nb, s));
        if (g.ns == 2) NCA_CUDA_OK(cudaMemsetAsync(G[0], 0, 2 * nc * sizeof(float), s));
        const bool chist = coarse_hist != nullptr && g.ns == 2;
        uint8_t* op_scratch = (uint8_t*)workspace + nca_align_up((size_t)((uint8_t*)(wsXc + 3 * nc) - (uint8_t*)workspace), 256);
        DyncaTc2Maps xm, gm_final, gm[2];
        NCA_CHECK_ARG(NCA_ALIGNED16(cond) && NCA_ALIGNED16(g_final), "cond / g_final must be 16-byte aligned");
        rc = dynca_tc2_make_maps(g, states, T + 1, chist ? coarse_hist : wsXc, chist ? T + 1 : 1, ncx, cond, &xm);
        if (rc) return rc;
        for (int p = 0; p < 2; ++p) { rc = dynca_tc2_make_gmaps(g, F[p], G[p], &gm[p]); if (rc) return rc; }
        if (g_final) { rc = dynca_tc2_make_gmaps(g, g_final, G[1], &gm_final); if (rc) return rc; }
        int p_in = 1;          // buffer index feeding the current step (step T-1 reads g_final or the zero buffer F[1], and G[1])
        for (int t = T - 1; t >= 0; --t) {
            FireMask fm = make_mask(d, g, masks, seed, t);
            fm.t = (uint32_t)(t0 + t);
            const bool from_final = (t == T - 1) && g_final != nullptr;
            const int p_out = 1 - p_in;
            float* gout = t == 0 ? gx0 : F[p_out];
            const float* tap = nullptr;
            if (ti >= 0 && tap_steps[ti] == t + 1) tap = g_taps[ti--];
            const uint8_t* op_t = op_hist ? (const uint8_t*)op_hist + (size_t)t * dynca_tc2_op_hist_bytes(g, 1) : nullptr;
            if (op_t == nullptr) {
                // no operand history for this step: the forward kernel records the operand of states[t] into the workspace and stops
                const float* xc_t = nullptr;
                if (g.ns == 2) {
                    if (chist) xc_t = coarse_hist + (size_t)t * ncx;
                    else { rc = dynca_bf16_coarsen(g, states + (size_t)t * n, wsXc, s); if (rc) return rc; xc_t = wsXc; }
                }
                rc = dynca_tc2_forward_step(g, wsB, &xm, t, states + (size_t)t * n, nullptr, chist ? t : 0, xc_t, nullptr, cond, fm, s, 0, op_scratch, 1);
                if (rc) return rc;
                op_t = op_scratch;
            }
            rc = dynca_tc2_backward_step(g, wsB, wsG, from_final ? &gm_final : &gm[p_in], from_final ? const_cast<float*>(g_final) : F[p_in], G[p_in],
                                         from_final ? 0 : 1, 1, tap, tap_c, tap_scale, gout, G[p_out], fm, s,
                                         t < T - 1 && op_hist != nullptr, op_t);
            if (rc) return rc;
            p_in = p_out;
        }
        if (g.ns == 2) { rc = dynca_tc2_add_coarse(g, G[p_in], gx0, s); if (rc) return rc; }
        return dynca_f32_unpack_grads(g, wsG, gw, s);
    }
    const float* gnext = g_final;   // NULL = zeros
    for (int t = T - 1; t >= 0; --t) {
        FireMask fm = make_mask(d, g, masks, seed, t);
        fm.t = (uint32_t)(t0 + t);
        float* gout = t == 0 ? gx0 : gbuf[t & 1];
        NCA_CUDA_OK(cudaMemsetAsync(gout, 0, nb, s));
        const float* tap = nullptr;   // gradient injected at states[t+1]
        if (ti >= 0 && tap_steps[ti] == t + 1) tap = g_taps[ti--];
        const float* xc_t = (coarse_hist && g.ns == 2) ? coarse_hist + (size_t)t * g.B * g.C * (g.H / 2) * (g.W / 2) : nullptr;
        rc = bf16 ? dynca_bf16_backward_step(g, wsB, wsXc, xc_t, wsG, states + (size_t)t * n, gnext, tap, tap_c, tap_scale, gout, cond, fm, s, x3)
                  : dynca_f32_backward_step(g, wsW, wsG, states + (size_t)t * n, gnext, tap, tap_c, tap_scale, gout, cond, fm, s);
        if (rc) return rc;
        gnext = gout;
    }
    return dynca_f32_unpack_grads(g, wsG, gw, s);
}

/* ---- ConditionedNCA: implemented in enc_f32.cu ---- */

}  // extern "C"
