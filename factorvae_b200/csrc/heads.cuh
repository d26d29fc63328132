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
    float *c1_mu, *c1_sg;             // [B][K]  sum_i beta_ik dmu_y_i, sum_i beta_ik^2 dvar_i
    // tensor-core backward sweep (heads_tc.cu): per-date vectors handed from the vector kernel to the sweep kernel
    float *t_dyp;        // [B][M]     d loss / d y_p
    float *t_dps;        // [B][K][H]  dp_k = Wv_k^T dctx_k
    float *t_pdp;        // [B][K]     pooled_k . dp_k
    int *t_tile_ptr;     // [B+1]      prefix of 128-stock tiles per date
    void *t_b1, *t_b2;   // bf16 operand images of the stacked weight rows (date independent part)
};

struct HeadsParts {   // stand-alone sub-module calls (fvae_heads_parts): optional inputs / outputs of the per-date heads
    const float *z_mu, *z_sigma;      // [B][K]  factors handed to the decoder (FactorDecoder.forward's arguments) or NULL
    float *alpha_mu, *alpha_sigma;    // [S]     AlphaLayer.forward                       module.py:78-84
    float *beta;                      // [S][K]  BetaLayer.forward                        module.py:92-94
    float *context;                   // [B][K][H]  AttentionLayer.forward of every head  module.py:134-153
};

struct HeadsArgs {
    int S, B, H, K, M;
    const int* date_ptr;        // [B+1]
    const float* e;             // [S][H]
    const float* y;             // [S] (NULL in predict mode)
    fvae_noise noise;
    uint32_t flags;
    int predict;                // 1: FactorVAE.prediction (prior factors into the decoder)
    int use_tc;                 // FVAE_PREC_BF16_TC: the backward sweep may run on tcgen05 (heads_tc.cu)
    fvae_outputs out;
    HeadsW w;
    HeadsSaved sv;
    HeadsParts parts;           // all NULL inside the ELBO step
};

#ifdef __CUDACC__
// The Philox step counter of this launch: a kernel argument, or -- for steps captured in a CUDA graph, whose arguments are frozen
// at capture -- a device word the graph itself advances (fvae_noise.step_dev).  Read once at the top of a kernel.
__device__ __forceinline__ uint64_t noise_step(const HeadsArgs& a) {
    return a.noise.step_dev ? *reinterpret_cast<const volatile uint64_t*>(a.noise.step_dev) : a.noise.step;
}
__device__ __forceinline__ float keep_factor(const HeadsArgs& a, uint64_t step, int unit, int k) {
    // dropout on the attention scores (module.py:144): kept -> 1/0.9, dropped -> 0; eval -> 1
    if (!(a.flags & FVAE_FLAG_TRAIN)) return 1.f;
    bool keep;
    if (a.noise.keep_mask) keep = a.noise.keep_mask[size_t(unit) * a.K + k] != 0;
    else keep = philox_keep(a.noise.seed, step, a.noise.unit_base + unit, k);
    return keep ? kKeepScale : 0.f;
}
__device__ __forceinline__ float eps_of(const HeadsArgs& a, uint64_t step, int unit) {
    if (a.noise.eps) return a.noise.eps[unit];
    return philox_normal(a.noise.seed, step, a.noise.unit_base + unit);
}
// relu that propagates NaN like torch (fmaxf would swallow it)
__device__ __forceinline__ float relu_nan(float s) { return (s > 0.f || s != s) ? s : 0.f; }
#endif

// tensor-core backward sweep (heads_tc.cu): column layout of the stacked weight rows, padded to 8-column groups
struct TcCols {
    int Kp, Hp8;                 // K, H padded to 8
    int c_att, c_beta, c_alpha;  // static groups: [0,M) encoder | attention scores | beta | alpha hidden
    int NS;                      // static columns, padded to 16
    int c_atta;                  // = NS: 32 per-date columns (attention weights; rows = dp_k of the date)
    int NZ;                      // NS + 32
};
__host__ __device__ inline TcCols tc_cols(int H, int K, int M) {
    TcCols c;
    c.Kp = (K + 7) & ~7; c.Hp8 = (H + 7) & ~7;
    c.c_att = M; c.c_beta = M + c.Kp; c.c_alpha = M + 2 * c.Kp;
    c.NS = (M + 2 * c.Kp + c.Hp8 + 15) & ~15;
    c.c_atta = c.NS;
    c.NZ = c.NS + 32;
    return c;
}
// M == 128: the encoder rows fill exactly the first 128-column block (one UMMA M block of the weight-gradient GEMM)
inline bool heads_tc_supported(int H, int K, int M) { return M == 128 && H <= 31 && K <= 32; }
int64_t heads_tc_image_bytes(int H, int K, int M, int which);     // which: 1 = forward-product image, 2 = dE image
int heads_tc_prep(const HeadsArgs& a, cudaStream_t stream);        // builds the images (after heads_prep)
int heads_tc_forward(const HeadsArgs& a, cudaStream_t stream);       // forward of the heads (images must be built)
int heads_tc_sweep(const HeadsArgs& a, const HeadsG& g, float* dE, cudaStream_t stream);

// all launchers are asynchronous on `stream` and return a cudaError_t as int
int heads_prep(const HeadsArgs& a, bool zero_grad_acc, cudaStream_t stream);
int heads_forward(const HeadsArgs& a, cudaStream_t stream);
int heads_backward(const HeadsArgs& a, const HeadsG& g, float* dE /*[S][H]*/, cudaStream_t stream);
int heads_post(const HeadsArgs& a, const HeadsG& g, cudaStream_t stream);
int loss_reduce(const float* date_loss, int B, float* loss, cudaStream_t stream);

}  // namespace fvae
