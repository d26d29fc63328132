// Cross-sectional "heads" of the ELBO step: FactorEncoder, FactorPredictor (collapsed K-head
// attention), FactorDecoder (alpha/beta), reparameterisation, MSE + KL -- forward and backward.
// Interface between the C ABI (fvae_abi.cu) and the kernels (heads.cu).
#pragma once
#include "fvae_common.cuh"

namespace fvae {

struct HeadsW {   // parameter pointers (reference names in comments)
    const float *Wp, *bp;              // factor_encoder.linear            (M,H),(M)
    const float *Wmu, *bmu;            // factor_encoder.linear_mu         (K,M),(K)
    const float *Wsig, *bsig;          // factor_encoder.linear_sigma      (K,M),(K)
    const float *Wa, *ba;              // alpha_layer.linear1              (H,H),(H)
    const float *wam, *bam;            // alpha_layer.mu_layer             (H),(1)
    const float *was, *bas;            // alpha_layer.sigma_layer          (H),(1)
    const float *Wb, *bb;              // beta_layer.linear1               (K,H),(K)
    const float *q, *Wk, *bk, *Wv, *bv;  // attention_layers.{k}.*         stacked (K,H),(K,H,H),(K,H),(K,H,H),(K,H)
    const float *Wl, *bl;              // factor_predictor.linear          (H,H),(H)
    const float *wpm, *bpm;            // factor_predictor.mu_layer        (H),(1)
    const float *wps, *bps;            // factor_predictor.sigma_layer     (H),(1)
};

struct HeadsG {   // gradient pointers, same order
    float *Wp, *bp, *Wmu, *bmu, *Wsig, *bsig, *Wa, *ba, *wam, *bam, *was, *bas, *Wb, *bb;
    float *q, *Wk, *bk, *Wv, *bv, *Wl, *bl, *wpm, *bpm, *wps, *bps;
};

struct HeadsSaved {   // per-step state kept in the workspace between forward and backward
    float *G;            // [K][H]   Wk_k^T q_k
    float *cvec;         // [K]      q_k . bk_k
    float *dG;           // [K][H]   backward accumulators
    float *dc;           // [K]
    float *enc_m, *enc_l, *yp;        // [B][M]  column-softmax max / sum, portfolio returns
    float *att_m, *att_l;             // [B][K]
    float *pooled;                    // [B][K][H]  sum_i a_ik e_i
    float *ctx;                       // [B][K][H]
    float *hm_pre;                    // [B][K][H]
    float *pre_sg_post, *pre_sg_prior;  // [B][K]  softplus inputs
    int *bad;                         // [B][K]  NaN/Inf guard tripped (module.py:149)
    int *clamp_post, *clamp_prior;    // [B][K]  sigma == 0 -> 1e-6 fired (module.py:117 / :265)
};

struct HeadsArgs {
    int S, B, H, K, M;
    const int* date_ptr;        // [B+1]
    const float* e;             // [S][H]
    const float* y;             // [S] (NULL in predict mode)
    fvae_noise noise;
    uint32_t flags;
    int predict;                // 1: FactorVAE.prediction (prior factors into the decoder)
    fvae_outputs out;
    HeadsW w;
    HeadsSaved sv;
};

// all launchers are asynchronous on `stream` and return a cudaError_t as int
int heads_prep(const HeadsArgs& a, bool zero_grad_acc, cudaStream_t stream);
int heads_forward(const HeadsArgs& a, cudaStream_t stream);
int heads_backward(const HeadsArgs& a, const HeadsG& g, float* dE /*[S][H]*/, cudaStream_t stream);
int heads_post(const HeadsArgs& a, const HeadsG& g, cudaStream_t stream);
int loss_reduce(const float* date_loss, int B, float* loss, cudaStream_t stream);

}  // namespace fvae
