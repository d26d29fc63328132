// FeatureExtractor, FVAE_PREC_BF16_TC: bf16 operands on tcgen05 tensor cores (sm_100a).
// (placeholder: filled in after the fp32 path is parity-green on the GPU)
#include "fe.cuh"

namespace fvae {
int64_t fe_tc_workspace_bytes(const FeDims&) { return 256; }
int fe_tc_supported(const FeDims&) { return FVAE_ERR_UNSUPPORTED; }
int fe_tc_forward(const FeDims&, const fvae_panel&, const FeW&, float*, void*, cudaStream_t) { return FVAE_ERR_UNSUPPORTED; }
int fe_tc_backward(const FeDims&, const fvae_panel&, const FeW&, const FeG&, const float*, void*, cudaStream_t) {
    return FVAE_ERR_UNSUPPORTED;
}
}  // namespace fvae
