// Internal launcher prototypes shared between the translation units of libnca_b200.so (not part of the ABI).
#pragma once
#include "dynca_common.cuh"

// dynca_f32.cu
size_t dynca_f32_weight_floats(const DyncaGeom& g);
size_t dynca_f32_grad_floats(const DyncaGeom& g);
int dynca_f32_prep_weights(const DyncaGeom& g, const NcaDyncaWeights* w, float* ws, cudaStream_t s);
int dynca_f32_forward_step(const DyncaGeom& g, const float* wsW, const float* x_in, float* x_out, const float* cond,
                           const FireMask& fm, cudaStream_t s);
int dynca_f32_perceive(const DyncaGeom& g, const float* x, const float* cond, float* z, cudaStream_t s);
int dynca_f32_backward_step(const DyncaGeom& g, const float* wsW, float* wsG, const float* x_in, const float* g_next,
                            const float* g_tap, int tap_c, float tap_scale, float* g_out, const float* cond,
                            const FireMask& fm, cudaStream_t s);
int dynca_f32_unpack_grads(const DyncaGeom& g, const float* wsG, const NcaDyncaWeightGrads* gw, cudaStream_t s);
int nca_edge_extract_launch(int B, int H, int W, const float* img, int tanh_transform, float* out, cudaStream_t s);
int nca_philox_mask_launch(int B, int H, int W, float rate, int enc, uint64_t seed, int t0, const uint32_t* t0_dev, int T, float* out, cudaStream_t s);

// dynca_bf16.cu (tcgen05 path)
// x3 = NCA_PREC_F16X3: hi + lo operand images, three MMAs per product
size_t dynca_bf16_weight_bytes(const DyncaGeom& g, bool x3 = false);
bool dynca_bf16_supported(const DyncaGeom& g, bool x3);
int dynca_bf16_prep_weights(const DyncaGeom& g, const NcaDyncaWeights* w, void* ws, cudaStream_t s, bool x3 = false);
int dynca_bf16_forward_step(const DyncaGeom& g, const void* ws, float* xc, const float* x_in, float* x_out, const float* cond,
                            const FireMask& fm, cudaStream_t s, bool x3 = false);
size_t dynca_bf16_coarse_floats(const DyncaGeom& g);
size_t dynca_bf16_bwd_weight_bytes(const DyncaGeom& g, bool x3 = false);
bool dynca_bf16_bwd_supported(const DyncaGeom& g, bool x3 = false);
int dynca_bf16_prep_bwd_weights(const DyncaGeom& g, const NcaDyncaWeights* w, void* ws, cudaStream_t s, bool x3 = false);
int dynca_bf16_backward_step(const DyncaGeom& g, const void* ws, float* xc, const float* xc_ready, float* wsG, const float* x_in, const float* g_next,
                             const float* g_tap, int tap_c, float tap_scale, float* g_out, const float* cond,
                             const FireMask& fm, cudaStream_t s, bool x3 = false);

// dynca_tc2.cu (tcgen05 path, 8x16 tiles + TMA; shapes with W % 4 == 0, fc % 32 == 0)
struct DyncaTc2Maps { alignas(64) unsigned char x[128]; alignas(64) unsigned char xc[128]; alignas(64) unsigned char cond[128]; };   // CUtensorMaps
bool dynca_tc2_supported(const DyncaGeom& g);
size_t dynca_tc2_weight_bytes(const DyncaGeom& g);
int dynca_tc2_prep_weights(const DyncaGeom& g, const NcaDyncaWeights* w, void* ws, cudaStream_t s);
int dynca_tc2_make_maps(const DyncaGeom& g, const float* states, int slots, const float* coarse, int cslots, size_t cslot_floats,
                        const float* cond, DyncaTc2Maps* m);
// ops_only: record the perception operand of the step in op_out and do nothing else (the BPTT's recompute path)
int dynca_tc2_forward_step(const DyncaGeom& g, const void* ws, const DyncaTc2Maps* m, int slot_in, const float* x_in, float* x_out,
                           int cslot_in, const float* xc_in, float* xc_out, const float* cond, const FireMask& fm, cudaStream_t s, int pdl,
                           uint8_t* op_out, int ops_only = 0);
size_t dynca_tc2_op_hist_bytes(const DyncaGeom& g, int T);
int dynca_bf16_coarsen(const DyncaGeom& g, const float* x, float* xc, cudaStream_t s);

// dynca_tc3_bwd.cu (tcgen05 BPTT, 8x16 tiles, transposed perception gradient + register stencils; needs the recorded operand)
bool dynca_tc2_bwd_supported(const DyncaGeom& g);
size_t dynca_tc2_bwd_weight_bytes(const DyncaGeom& g);
int dynca_tc2_prep_bwd_weights(const DyncaGeom& g, const NcaDyncaWeights* w, void* ws, cudaStream_t s);
int dynca_tc2_make_gmaps(const DyncaGeom& g, const float* gfine, const float* gcoarse, DyncaTc2Maps* m);
int dynca_tc2_add_coarse(const DyncaGeom& g, const float* gc, float* gx, cudaStream_t s);
int dynca_tc2_backward_step(const DyncaGeom& g, const void* ws, float* wsG, const DyncaTc2Maps* gm, float* g_in, float* gc_in, int zero_in,
                            int zero_cin, const float* g_tap, int tap_c, float tap_scale, float* g_out, float* gc_out, const FireMask& fm,
                            cudaStream_t s, int pdl, const uint8_t* op_in);

// enc_tc.cu (ConditionedNCA forward on tcgen05)
struct EncTcMaps { alignas(64) unsigned char x[128]; alignas(64) unsigned char l[128]; alignas(64) unsigned char g[128]; };
bool enc_tc_supported(const NcaEncDesc* d);
size_t enc_tc_weight_bytes(const NcaEncDesc* d);
int enc_tc_prep_weights(const NcaEncDesc* d, const NcaEncWeights* w, void* ws, cudaStream_t s);
int enc_tc_make_maps(const NcaEncDesc* d, const float* states, int slots, const float* goal, EncTcMaps* m);
int enc_tc_forward_step(const NcaEncDesc* d, const NcaEncWeights* w, const void* ws, const EncTcMaps* m, int slot_in, float* x1,
                        const FireMask& fm, cudaStream_t s, int pdl);
bool enc_tc_bwd_supported(const NcaEncDesc* d);
int enc_tc_make_gmap(const NcaEncDesc* d, const float* gnext, EncTcMaps* m);
int enc_tc_backward_step(const NcaEncDesc* d, const NcaEncWeights* w, const void* ws, const EncTcMaps* m, const EncTcMaps* gm,
                         int slot_in, const uint8_t* life, float* g_out, float* g_goal, const NcaEncWeightGrads* gw,
                         const FireMask& fm, cudaStream_t s);
