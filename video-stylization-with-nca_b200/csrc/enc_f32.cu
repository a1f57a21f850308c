// ConditionedNCA step (EncoderConditioning/nca.py:152-209), fp32 path.  (placeholder entry points)
#include "nca_internal.h"

extern "C" {
size_t nca_enc_workspace_bytes(const NcaEncDesc* d, int32_t backward) { (void)d; (void)backward; return 0; }
int nca_enc_forward(const NcaEncDesc*, const NcaEncWeights*, const float*, const float*, uint64_t, int32_t, int32_t,
                    int32_t, float*, uint8_t*, void*, size_t, void*) {
    nca_set_error("ConditionedNCA kernels are not built into this library yet");
    return NCA_ERR_UNSUPPORTED;
}
int nca_enc_backward(const NcaEncDesc*, const NcaEncWeights*, const float*, const float*, uint64_t, int32_t, int32_t,
                     const float*, const uint8_t*, const float*, float*, float*, const NcaEncWeightGrads*, void*, size_t,
                     void*) {
    nca_set_error("ConditionedNCA kernels are not built into this library yet");
    return NCA_ERR_UNSUPPORTED;
}
}
