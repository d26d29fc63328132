// FeatureExtractor (reference module.py:22-31): LayerNorm -> Linear -> LeakyReLU -> GRU -> h_T.
// Interface between the C ABI and the two implementations:
//   fe_f32.cu : FVAE_PREC_FP32     CUDA-core FFMA, fp32 everywhere (tight parity mode)
//   fe_tc.cu  : FVAE_PREC_BF16_TC  bf16 operands on tcgen05 tensor cores, fp32 accumulate
#pragma once
#include "fvae_common.cuh"

namespace fvae {

struct FeW {   // feature_extractor.{normalize,linear,gru.*}
    const float *ln_w, *ln_b;   // (C),(C)
    const float *W1, *b1;       // (C,C),(C)
    const float *Wih, *Whh;     // (3H,C),(3H,H)   row blocks [r; z; n]
    const float *bih, *bhh;     // (3H),(3H)
};
struct FeG { float *ln_w, *ln_b, *W1, *b1, *Wih, *Whh, *bih, *bhh; };

struct FeDims { int S, T, C, H; };

// fp32 workspace (bytes) beyond e / dE: gi [S][T][3H], hall [S][T][H], dgh [S][T][3H], row-chunk scratch
int64_t fe_f32_workspace_bytes(const FeDims& d);
// forward: writes e[S][H] (= h_T) and keeps gi / hall in `ws` for backward
int fe_f32_forward(const FeDims& d, const fvae_panel& x, const FeW& w, float* e, void* ws, cudaStream_t stream);
// backward: consumes dE[S][H]; ACCUMULATES into the FeatureExtractor gradient sections
int fe_f32_backward(const FeDims& d, const fvae_panel& x, const FeW& w, const FeG& g, const float* dE, void* ws,
                    cudaStream_t stream);

// tensor-core path (same contract); prepared bf16 weights live in its workspace
int64_t fe_tc_workspace_bytes(const FeDims& d);
int fe_tc_supported(const FeDims& d);
int fe_tc_forward(const FeDims& d, const fvae_panel& x, const FeW& w, float* e, void* ws, cudaStream_t stream);
int fe_tc_front_only(const FeDims& d, const fvae_panel& x, void* ws, cudaStream_t stream);   // diagnostics: K1 alone
int fe_tc_backward(const FeDims& d, const fvae_panel& x, const FeW& w, const FeG& g, const float* dE, void* ws,
                   cudaStream_t stream);

}  // namespace fvae
