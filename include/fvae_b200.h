/* fvae_b200.h -- C ABI of the B200-native FactorVAE ELBO-step hot path.
 *
 * Drop-in boundary.  The reference has no FFI: its boundary is the
 * Python call `FactorVAE.forward(x, returns)` (reference module.py:250-270) followed by
 * `loss.backward()` (train_model.py:29), and `FactorVAE.prediction(x)` (module.py:273-278).
 * Every entry point below replaces the PyTorch/ATen arithmetic behind one of those calls:
 *
 *   fvae_elbo_forward   <- FactorVAE.forward            module.py:250-270
 *                          (FeatureExtractor :22-31, FactorEncoder :52-67, FactorDecoder :107-123,
 *                           FactorPredictor :169-188 / AttentionLayer :134-153, KL :242-248)
 *   fvae_elbo_backward  <- loss.backward()              train_model.py:29 (autograd of the above)
 *   fvae_predict        <- FactorVAE.prediction         module.py:273-278
 *   fvae_fe_forward     <- FeatureExtractor.forward     module.py:22-31
 *   fvae_fe_backward    <- autograd of FeatureExtractor.forward
 *   fvae_heads_parts    <- FactorEncoder / AlphaLayer / BetaLayer / FactorDecoder / AttentionLayer / FactorPredictor .forward
 *                          called on their own   module.py:52-67, :78-84, :92-94, :107-123, :134-153, :169-188
 *   fvae_param_*        <- the nn.Parameter inventory   module.py:17-20,37-41,72-75,90,129-131,163-166
 *
 * Conventions
 *   - plain C, no exceptions, no allocation, no global mutable state;
 *   - all pointers are DEVICE pointers unless named host_*; sizes are explicit;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and re-entrant
 *     across streams / devices;
 *   - return value: 0 ok; <0 invalid argument (fvae_status_string); >0 a cudaError_t;
 *   - the caller owns every buffer, including the workspace (size from fvae_workspace_bytes).
 *
 * Units follow the reference's domain: a *date* is one cross-section (one reference batch,
 * dataset.py:207-238), a *stock* is one sequence (one row of that batch); S = total
 * date x stock units in the call, split into B dates by the CSR array date_ptr[B+1].
 */
#ifndef FVAE_B200_H
#define FVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVAE_ABI_VERSION 6

/* status codes (<0: argument errors) */
#define FVAE_OK 0
#define FVAE_ERR_NULL (-1)        /* a required pointer is NULL                         */
#define FVAE_ERR_SHAPE (-2)       /* non-positive or inconsistent S/B/T/C/H/K/M         */
#define FVAE_ERR_LIMIT (-3)       /* outside the supported range (C<=192, H<=64, ...)   */
#define FVAE_ERR_DTYPE (-4)       /* unknown dtype / precision enum                     */
#define FVAE_ERR_WORKSPACE (-5)   /* workspace too small or misaligned                  */
#define FVAE_ERR_NO_DEVICE (-6)   /* no CUDA device / wrong architecture                */
#define FVAE_ERR_UNSUPPORTED (-7) /* combination not implemented on this path           */

/* panel element type */
#define FVAE_F32 0
#define FVAE_BF16 1

/* arithmetic of the FeatureExtractor contractions */
#define FVAE_PREC_FP32 0     /* CUDA-core FFMA, fp32 everywhere: tight parity mode           */
#define FVAE_PREC_BF16_TC 1  /* bf16 operands on tcgen05 tensor cores, fp32 accumulate       */

/* flags */
#define FVAE_FLAG_TRAIN 1u        /* dropout on attention scores is active (module.py:144)  */
#define FVAE_FLAG_PHILOX 2u       /* eps / keep-masks come from the in-kernel Philox stream  */

typedef struct fvae_shape {
    int32_t S;   /* date x stock units (sequences) in this call                          */
    int32_t B;   /* dates                                                                */
    int32_t T;   /* look-back window (seq_len, main.py:98)                               */
    int32_t C;   /* features per row (num_latent, 158)                                   */
    int32_t H;   /* hidden_size                                                          */
    int32_t K;   /* num_factor                                                           */
    int32_t M;   /* num_portfolio                                                        */
} fvae_shape;

/* the feature panel x[S][T][C]; innermost stride is 1.  The reference's CPU path feeds a
 * non-contiguous view with row pitch 159 (train_model.py:18): pitches are explicit.
 * Resident-panel form (replaces the per-sample window gather of dataset.py:139-181): `data` is
 * the (date, instrument) row table [num_rows][row_pitch] kept on the device once, and
 * `row_index[s*T + t]` names the table row that is time step t of sequence s (what
 * TSDataSampler._get_indices returns after ffill+bfill; missing -> the all-NaN sentinel row).
 * The windows are then never materialised: the kernels read the T rows of a sequence in place. */
typedef struct fvae_panel {
    const void* data;
    int32_t dtype;        /* FVAE_F32 | FVAE_BF16 */
    int64_t seq_pitch;    /* elements between consecutive sequences (ignored with row_index) */
    int64_t row_pitch;    /* elements between consecutive time rows  */
    const int32_t* row_index;  /* NULL: dense windows.  Else [S][T] row numbers into `data` (0 <= r < num_rows) */
    int64_t num_rows;          /* rows of the table when row_index is given                                     */
} fvae_panel;

/* random inputs of the step.  Either explicit tensors (parity mode) or a Philox key. */
typedef struct fvae_noise {
    const float* eps;          /* [S]   N(0,1) draws for reparameterize (module.py:104) or NULL */
    const uint8_t* keep_mask;  /* [S][K] 1=keep for dropout on scores (module.py:144) or NULL   */
    uint64_t seed;             /* Philox key  (FVAE_FLAG_PHILOX)                                */
    uint64_t step;             /* Philox counter high word: training step                       */
    int64_t unit_base;         /* global index of unit 0 of this call (shard-invariant RNG)     */
    const uint64_t* step_dev;  /* NULL, or a DEVICE word holding the step counter: it overrides `step` and is read by the
                                  kernels at run time, so a step captured in a CUDA graph (whose kernel arguments are frozen at
                                  capture) draws fresh noise on every replay -- the graph itself advances the word        */
} fvae_noise;

typedef struct fvae_outputs {
    float* loss;        /* [1]    mean over dates of (mse + KL)   module.py:268              */
    float* date_loss;   /* [B]    per-date vae_loss                                           */
    float* yhat;        /* [S]    reconstruction sample           module.py:123               */
    float* mu_y;        /* [S]    module.py:120 (a temporary in the reference)                */
    float* sigma_y;     /* [S]    module.py:121                                               */
    float* mu_post;     /* [B][K] factor_mu                                                   */
    float* sigma_post;  /* [B][K] factor_sigma (after the :117 clamp)                         */
    float* mu_prior;    /* [B][K] pred_mu                                                     */
    float* sigma_prior; /* [B][K] pred_sigma (after the :265 clamp)                           */
} fvae_outputs;

/* ---- parameter inventory: one flat fp32 buffer, sections 16-byte aligned ---------------- */
enum fvae_param_section {
    FVAE_P_LN_W = 0, FVAE_P_LN_B, FVAE_P_W1, FVAE_P_B1, FVAE_P_WIH, FVAE_P_WHH, FVAE_P_BIH, FVAE_P_BHH,
    FVAE_P_ENC_W, FVAE_P_ENC_B, FVAE_P_ENC_MU_W, FVAE_P_ENC_MU_B, FVAE_P_ENC_SG_W, FVAE_P_ENC_SG_B,
    FVAE_P_AL_W, FVAE_P_AL_B, FVAE_P_AL_MU_W, FVAE_P_AL_MU_B, FVAE_P_AL_SG_W, FVAE_P_AL_SG_B,
    FVAE_P_BETA_W, FVAE_P_BETA_B,
    FVAE_P_ATT_Q, FVAE_P_ATT_KW, FVAE_P_ATT_KB, FVAE_P_ATT_VW, FVAE_P_ATT_VB,   /* stacked over the K heads */
    FVAE_P_PR_W, FVAE_P_PR_B, FVAE_P_PR_MU_W, FVAE_P_PR_MU_B, FVAE_P_PR_SG_W, FVAE_P_PR_SG_B,
    FVAE_P_NUM_SECTIONS
};

int fvae_abi_version(void);
/* diagnostics: kernels launched by this library since load (bench.py's gpu_launches) */
uint64_t fvae_debug_launch_count(void);
const char* fvae_status_string(int status);

/* offsets[FVAE_P_NUM_SECTIONS + 1] in floats; the last entry is the flat buffer length. */
int fvae_param_offsets(int32_t C, int32_t H, int32_t K, int32_t M, int64_t* host_offsets);
int64_t fvae_param_count(int32_t C, int32_t H, int32_t K, int32_t M);

/* bytes of caller-provided workspace (saved activations of the step + scratch). */
int64_t fvae_workspace_bytes(const fvae_shape* shape, int32_t precision);

/* forward of one ELBO step over B dates.  Saves what backward needs in `workspace`. */
int fvae_elbo_forward(const fvae_shape* shape, const fvae_panel* x, const float* y /*[S]*/,
                      const int32_t* date_ptr /*[B+1]*/, const float* params, const fvae_noise* noise,
                      uint32_t flags, int32_t precision, const fvae_outputs* out,
                      void* workspace, int64_t workspace_bytes, void* stream);

/* backward of the step just run by fvae_elbo_forward with the same arguments and workspace.
 * Writes d(loss)/d(params) into grad[param_count] (overwrites; the layout is the params'). */
int fvae_elbo_backward(const fvae_shape* shape, const fvae_panel* x, const float* y,
                       const int32_t* date_ptr, const float* params, const fvae_noise* noise,
                       uint32_t flags, int32_t precision, const fvae_outputs* out, float* grad,
                       void* workspace, int64_t workspace_bytes, void* stream);

/* FactorVAE.prediction: prior factors into the decoder; y/date_loss/post outputs unused. */
int fvae_predict(const fvae_shape* shape, const fvae_panel* x, const int32_t* date_ptr,
                 const float* params, const fvae_noise* noise, uint32_t flags, int32_t precision,
                 const fvae_outputs* out, void* workspace, int64_t workspace_bytes, void* stream);

/* FeatureExtractor alone: e[S][H] = h_T; backward takes d(loss)/de and ACCUMULATES the eight
 * FeatureExtractor sections of `grad` (the caller zeroes grad). */
int fvae_fe_forward(const fvae_shape* shape, const fvae_panel* x, const float* params, int32_t precision,
                    float* e, void* workspace, int64_t workspace_bytes, void* stream);
int fvae_fe_backward(const fvae_shape* shape, const fvae_panel* x, const float* params, int32_t precision,
                     const float* de, float* grad, void* workspace, int64_t workspace_bytes, void* stream);

/* Stand-alone calls of the per-date sub-modules on caller-supplied stock latents e[S][H] (what FeatureExtractor.forward
 * returns), replacing the PyTorch arithmetic of
 *     FactorEncoder.forward      module.py:52-67    -> out->mu_post, out->sigma_post            (needs y)
 *     FactorPredictor.forward    module.py:169-188  -> out->mu_prior, out->sigma_prior
 *     AttentionLayer.forward     module.py:134-153  -> parts->context [B][K][H] (zeros where the NaN/Inf guard :149 tripped)
 *     AlphaLayer.forward         module.py:78-84    -> parts->alpha_mu, parts->alpha_sigma [S]
 *     BetaLayer.forward          module.py:92-94    -> parts->beta [S][K]
 *     FactorDecoder.forward      module.py:107-123  -> out->yhat (sample), out->mu_y, out->sigma_y from the factors
 *                                                      parts->z_mu / z_sigma [B][K] (sigma == 0 -> 1e-6 as :117)
 * in ONE launch of the fp32 heads kernel per call.  y == NULL: the encoder and the loss are skipped (out->loss, date_loss,
 * mu_post, sigma_post may be NULL) and, without z_mu, the decoder is fed the prior like FactorVAE.prediction; y != NULL: the
 * whole per-date forward of FactorVAE.forward runs from e.  FVAE_FLAG_TRAIN is honoured (dropout on the attention scores).
 * shape->T and shape->C only size the parameter layout and the workspace (fvae_workspace_bytes(shape, FVAE_PREC_FP32)).
 * Forward only: the gradient of the heads exists inside fvae_elbo_backward. */
typedef struct fvae_parts {
    const float* z_mu;      /* [B][K] or NULL */
    const float* z_sigma;   /* [B][K] or NULL (both or neither) */
    float* alpha_mu;        /* [S] or NULL */
    float* alpha_sigma;     /* [S] or NULL (both or neither) */
    float* beta;            /* [S][K] or NULL */
    float* context;         /* [B][K][H] or NULL */
} fvae_parts;
int fvae_heads_parts(const fvae_shape* shape, const float* latent, const float* y, const int32_t* date_ptr, const float* params,
                     const fvae_noise* noise, uint32_t flags, const fvae_parts* parts, const fvae_outputs* out, void* workspace,
                     int64_t workspace_bytes, void* stream);

/* diagnostics: re-launch ONLY the dominant tensor-core kernel (front forward: LayerNorm -> GEMM -> LeakyReLU -> GEMM)
 * on a workspace that a previous fvae_elbo_forward(FVAE_PREC_BF16_TC) call prepared; bench.py times it with CUDA
 * events for the per-kernel roofline. */
int fvae_debug_front_forward(const fvae_shape* shape, const fvae_panel* x, void* workspace, int64_t workspace_bytes,
                             void* stream);

/* diagnostics: the noise a FVAE_FLAG_PHILOX step with this key draws -- eps[S] (module.py:104) and keep_mask[S][K]
 * (1 = keep; dropout on the attention scores, module.py:132,144) for units unit_base .. unit_base+S-1, evaluated by the
 * same device functions the step's kernels call.  Parity tests replay a Philox step through the CPU oracle with it and
 * measure the keep rate; either output may be NULL. */
int fvae_debug_noise(uint64_t seed, uint64_t step, int64_t unit_base, int64_t S, int32_t K, float* eps,
                     uint8_t* keep_mask, void* stream);

/* device e[S][H] of the last forward on this workspace (for tests / diagnostics). */
const float* fvae_workspace_latent(const fvae_shape* shape, int32_t precision, const void* workspace);

/* ---- resident panel: window index + gather (replaces dataset.py:139-181 TSDataSampler._get_indices /
 *      __getitem__ and the label slice of train_model.py:17-22) ---------------------------------------
 * idx_mat [D][I]: row number of (date d, instrument j) in the row table, -1 where the instrument has
 * no row that date (TSDataSampler.build_index, dataset.py:128-137).  For sample s = (sample_date[s],
 * sample_inst[s]) writes row_index[s][t] = idx_mat[date - T + 1 + t][inst] (dates before the first
 * one count as missing, dataset.py:141-143), then fill_mode FVAE_FILL_NONE | _FFILL | _FFILL_BFILL
 * (np_ffill, dataset.py:24-39,145-148); what is still missing becomes `nan_row` (the all-NaN
 * sentinel row, dataset.py:81-84,172).  If `label` (one value per table row) and `y` are given,
 * y[s] = label[row_index[s][T-1]] (the reference's returns[:, -1]).                                */
#define FVAE_FILL_NONE 0
#define FVAE_FILL_FFILL 1
#define FVAE_FILL_FFILL_BFILL 2
int fvae_window_index(const int32_t* idx_mat, int32_t D, int32_t I, const int32_t* sample_date,
                      const int32_t* sample_inst, int64_t S, int32_t T, int32_t fill_mode, int32_t nan_row,
                      int32_t* row_index, const float* label, float* y, void* stream);
/* materialise out[S][T][C] (contiguous, out_dtype FVAE_F32 | FVAE_BF16) from a resident panel -- for
 * callers that still want the reference's window tensor (DataLoader batch, train_model.py:17-19).   */
int fvae_gather_windows(const fvae_panel* panel, int64_t S, int32_t T, int32_t C, void* out, int32_t out_dtype,
                        void* stream);

/* ---- fused optimizer step over the flat buffers (replaces optimizer.step() of main.py:60 /
 *      train_model.py:30: torch.optim.Adam, single-tensor arithmetic, amsgrad off).  `step` is 1-based.
 *      grad_scale multiplies the gradient first (1.0, or 1/world after a SUM all-reduce).  All four
 *      buffers 16-byte aligned, n = fvae_param_count().  The per-batch CosineAnnealingLR of
 *      main.py:61 is a host formula: lr_t = eta_min + (lr0 - eta_min) * (1 + cos(pi t / T_max)) / 2.  */
int fvae_adam_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                   void* stream);

/* ---- the step's single collective, as one kernel over NVLink peer memory (the reference has no data parallelism:
 *      no counterpart).  Each rank allocates a communication buffer (fvae_p2p_alloc: cudaMalloc + IPC handle), the ranks
 *      exchange the 64-byte handles (host side, e.g. torch.distributed.all_gather_object) and map each other's buffers
 *      (fvae_p2p_open).  fvae_p2p_allreduce: inout[n] <- scale * sum_ranks inout[n], one launch: push to every peer's slot,
 *      publish per-chunk flags, wait for the peers' chunks, sum in rank order (bit-identical on all ranks).  `epoch` = 1, 2,
 *      3, ... the same on every rank; `peer_bases` is a DEVICE array of the world's buffer addresses as mapped here.        */
int64_t fvae_p2p_buffer_bytes(int64_t n, int32_t world);
int fvae_p2p_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle64);
int fvae_p2p_open(const unsigned char* handle64, void** dev_ptr);
int fvae_p2p_close(void* dev_ptr);
int fvae_p2p_free(void* dev_ptr);
int fvae_p2p_allreduce(float* inout, int64_t n, void* const* peer_bases, int32_t world, int32_t rank, uint32_t epoch, float scale,
                       int64_t n_alloc, void* stream);

/* ---- evaluation metric: per-date Spearman rank correlation of predictions and labels (replaces the
 *      per-date pandas .rank() + scipy.stats.spearmanr loop of utils.py:113-129).  ric[d] for the dates of
 *      date_ptr (CSR, B+1); average ranks for ties; NaN input, < 2 stocks or a constant column -> NaN.
 *      max_per_date >= the largest date (<= 4096).  RankIC = mean(ric), RankIC_IR = mean / std (ddof 0). */
int fvae_rank_ic(const float* pred, const float* label, const int32_t* date_ptr, int32_t B, int32_t max_per_date,
                 float* ric, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FVAE_B200_H */
