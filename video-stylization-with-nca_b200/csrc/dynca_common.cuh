// DyNCA geometry shared by the fp32 (CUDA-core) and bf16 (tcgen05) kernels.
#pragma once
#include "nca_common.cuh"

// Padded-weight layout used by the fp32 path (built per call by dynca_prep_weights in the caller's workspace).
// The first-layer bias is folded into W1 through a constant-1 perception row (k = P), so w1 and b1 gradients
// come out of the same outer-product accumulation.
//   W1p [Ppad][FCpad]: W1p[k][j] = w1[j][k] (k<P, j<fc); W1p[P][j] = b1[j]; else 0
//   W2p [FCpad][CP]  : W2p[j][c] = w2[c][j];   W2t [CP][FCpad] : W2t[c][j] = w2[c][j];   b2p [CP]
struct DyncaGeom {
    int B, C, H, W, fc, cc, P, cond_kind, pad, ns;
    int Ppad, FCpad, CP;
    float s0;              // 1 / n_scales
    float cpe_oh, cpe_ow;  // (float)(0.5/H), (float)(0.5/W)   (dynca.py:192-193)
};

static inline int dynca_ppad(int P) { return (P + 1 + 3) / 4 * 4; }
static inline int dynca_fcpad(int fc) { return (fc + 7) / 8 * 8; }

static inline int dynca_make_geom(const NcaDyncaDesc* d, DyncaGeom* g) {
    NCA_CHECK_ARG(d != nullptr, "desc is NULL");
    NCA_CHECK_ARG(d->B > 0 && d->C > 0 && d->H > 0 && d->W > 0 && d->fc > 0, "B,C,H,W,fc must be positive");
    NCA_CHECK_ARG(d->C <= 16, "C=%d > 16 is not supported", d->C);
    NCA_CHECK_ARG(d->fc <= 128, "fc=%d > 128 is not supported", d->fc);
    NCA_CHECK_ARG(d->pad_mode >= 0 && d->pad_mode <= 3, "bad pad_mode %d", d->pad_mode);
    NCA_CHECK_ARG(d->n_scales == 1 || d->n_scales == 2, "n_scales must be 1 ([0]) or 2 ([0,1]), got %d", d->n_scales);
    NCA_CHECK_ARG(d->cond_kind >= 0 && d->cond_kind <= 2, "bad cond_kind %d", d->cond_kind);
    int cc = d->cond_kind == NCA_COND_NONE ? 0 : (d->cond_kind == NCA_COND_CPE ? 2 : d->cc);
    NCA_CHECK_ARG(cc >= 0 && cc <= 16, "cc=%d out of range", cc);
    NCA_CHECK_ARG(d->cond_kind != NCA_COND_TENSOR || cc > 0, "cond tensor needs cc > 0");
    if (d->n_scales == 2) NCA_CHECK_ARG(d->H % 2 == 0 && d->W % 2 == 0, "multi-scale perception needs even H, W");
    if (d->pad_mode == NCA_PAD_REFLECT)
        NCA_CHECK_ARG(d->H / d->n_scales >= 2 && d->W / d->n_scales >= 2, "reflect padding needs >= 2 pixels per axis");
    NCA_CHECK_ARG(d->mask_mode == NCA_MASK_SUPPLIED || d->mask_mode == NCA_MASK_PHILOX, "bad mask_mode");
    NCA_CHECK_ARG((long long)d->H * d->W < (1ll << 31), "H*W too large");
    g->B = d->B; g->C = d->C; g->H = d->H; g->W = d->W; g->fc = d->fc; g->cc = cc;
    g->P = 4 * d->C + cc; g->cond_kind = d->cond_kind; g->pad = d->pad_mode; g->ns = d->n_scales;
    g->Ppad = dynca_ppad(g->P); g->FCpad = dynca_fcpad(d->fc); g->CP = 16;
    g->s0 = 1.0f / (float)d->n_scales;
    g->cpe_oh = (float)(0.5 / (double)d->H); g->cpe_ow = (float)(0.5 / (double)d->W);
    return NCA_OK;
}

// mask description handed to the step kernels (already offset to the step)
struct FireMask {
    const float* supplied;   // [B,1,H,W] for this step, or nullptr -> Philox
    uint32_t k0, k1, t;      // Philox key and step counter
    unsigned long long thr;  // 33-bit threshold
};

__device__ __forceinline__ float dynca_fire(const FireMask& m, int b, int y, int x, int H, int W) {
    if (m.supplied) return m.supplied[((size_t)b * H + y) * W + x];
    return nca_fire(nca_philox_word((uint32_t)(y * W + x), (uint32_t)b, m.t, m.k0, m.k1), m.thr, 0);
}

// CPE2D value (ExtraChannels/models/dynca.py:188-196): 2*((i/n - 0.5) + 0.5/n)
__device__ __forceinline__ float dynca_cpe(int i, int n, float off) {
    return 2.0f * ((__fdiv_rn((float)i, (float)n) - 0.5f) + off);
}

// bilinear x2 (align_corners=False, edge clamped) source description for fine index y on a coarse axis of nc:
// value = w0 * src[base] + w1 * src[base+1]; when the second tap is clamped away w1 = 0 (base+1 stays addressable
// in the padded coarse tile).
__device__ __forceinline__ void dynca_up_taps(int y, int nc, int& base, float& w0, float& w1) {
    int Q = y >> 1;
    if ((y & 1) == 0) {
        if (Q == 0) { base = 0; w0 = 1.0f; w1 = 0.0f; }
        else { base = Q - 1; w0 = 0.25f; w1 = 0.75f; }
    } else {
        base = Q;
        if (Q == nc - 1) { w0 = 1.0f; w1 = 0.0f; }
        else { w0 = 0.75f; w1 = 0.25f; }
    }
}
// transposed: weight with which fine index y reads coarse index q
__device__ __forceinline__ float dynca_up_weight(int y, int q, int nc) {
    int base; float w0, w1;
    dynca_up_taps(y, nc, base, w0, w1);
    return (base == q ? w0 : 0.0f) + (base + 1 == q ? w1 : 0.0f);
}
