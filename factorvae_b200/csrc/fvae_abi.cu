// C ABI of the FactorVAE hot path (include/fvae_b200.h): argument validation, workspace carving
// and the kernel chain of one ELBO step.  No allocation, no global state, asynchronous on `stream`.
#include <cstdio>
#include <cstring>

#include "fe.cuh"
#include "heads.cuh"

using namespace fvae;

namespace fvae { unsigned long long g_launch_count = 0; }

namespace {

struct Workspace {
    float *e, *dE;
    HeadsSaved sv;
    void* fe_ws;
    int64_t fe_bytes;
    int64_t bytes;
};

int64_t round256(int64_t v) { return (v + 255) & ~int64_t(255); }

Workspace carve(const fvae_shape& s, int precision, void* base) {
    Workspace w;
    char* p = static_cast<char*>(base);
    auto take = [&](int64_t nbytes) { char* r = p; p += round256(nbytes); return r; };
    const int64_t S = s.S, B = s.B, H = s.H, K = s.K, M = s.M;
    w.e = reinterpret_cast<float*>(take(S * H * 4));
    w.dE = reinterpret_cast<float*>(take(S * H * 4));
    w.sv.G = reinterpret_cast<float*>(take(K * H * 4));
    w.sv.cvec = reinterpret_cast<float*>(take(K * 4));
    w.sv.dG = reinterpret_cast<float*>(take(K * H * 4));
    w.sv.dc = reinterpret_cast<float*>(take(K * 4));
    w.sv.enc_m = reinterpret_cast<float*>(take(B * M * 4));
    w.sv.enc_l = reinterpret_cast<float*>(take(B * M * 4));
    w.sv.yp = reinterpret_cast<float*>(take(B * M * 4));
    w.sv.att_m = reinterpret_cast<float*>(take(B * K * 4));
    w.sv.att_l = reinterpret_cast<float*>(take(B * K * 4));
    w.sv.pooled = reinterpret_cast<float*>(take(B * K * H * 4));
    w.sv.ctx = reinterpret_cast<float*>(take(B * K * H * 4));
    w.sv.hm_pre = reinterpret_cast<float*>(take(B * K * H * 4));
    w.sv.pre_sg_post = reinterpret_cast<float*>(take(B * K * 4));
    w.sv.pre_sg_prior = reinterpret_cast<float*>(take(B * K * 4));
    w.sv.bad = reinterpret_cast<int*>(take(B * K * 4));
    w.sv.clamp_post = reinterpret_cast<int*>(take(B * K * 4));
    w.sv.clamp_prior = reinterpret_cast<int*>(take(B * K * 4));
    w.sv.c1_mu = reinterpret_cast<float*>(take(B * K * 4));
    w.sv.c1_sg = reinterpret_cast<float*>(take(B * K * 4));
    w.sv.t_dyp = reinterpret_cast<float*>(take(B * M * 4));
    w.sv.t_dps = reinterpret_cast<float*>(take(B * K * H * 4));
    w.sv.t_pdp = reinterpret_cast<float*>(take(B * K * 4));
    w.sv.t_tile_ptr = reinterpret_cast<int*>(take((B + 1) * 4));
    w.sv.t_b1 = w.sv.t_b2 = nullptr;
    if (precision == FVAE_PREC_BF16_TC && heads_tc_supported(s.H, s.K, s.M)) {
        w.sv.t_b1 = take(heads_tc_image_bytes(s.H, s.K, s.M, 1));
        w.sv.t_b2 = take(heads_tc_image_bytes(s.H, s.K, s.M, 2));
    }
    const FeDims fd{s.S, s.T, s.C, s.H};
    w.fe_bytes = (precision == FVAE_PREC_BF16_TC) ? fe_tc_workspace_bytes(fd) : fe_f32_workspace_bytes(fd);
    w.fe_ws = take(w.fe_bytes);
    w.bytes = p - static_cast<char*>(base);
    return w;
}

int check_shape(const fvae_shape* s, int precision) {
    if (!s) return FVAE_ERR_NULL;
    if (s->S <= 0 || s->B <= 0 || s->T <= 0 || s->C <= 0 || s->H <= 0 || s->K <= 0 || s->M <= 0 || s->B > s->S)
        return FVAE_ERR_SHAPE;
    // K, M bounds: the per-date phases of the heads kernels deal one factor / portfolio per thread of a 256-thread CTA
    // (heads.cu: NSC = NT / K would be 0 for K > 256 and the posterior-path column sums would silently be zero)
    if (s->C > kMaxC || s->H > kMaxH || s->K > 256 || s->M > 1024) return FVAE_ERR_LIMIT;
    if (precision != FVAE_PREC_FP32 && precision != FVAE_PREC_BF16_TC) return FVAE_ERR_DTYPE;
    if (precision == FVAE_PREC_BF16_TC) {
        const FeDims fd{s->S, s->T, s->C, s->H};
        const int rc = fe_tc_supported(fd);
        if (rc != 0) return rc;
    }
    return FVAE_OK;
}

int check_panel(const fvae_panel* x, const fvae_shape* s) {
    if (!x || !x->data) return FVAE_ERR_NULL;
    if (x->dtype != FVAE_F32 && x->dtype != FVAE_BF16) return FVAE_ERR_DTYPE;
    if (x->row_index) {                                      // resident panel: rows named by the index
        if (x->row_pitch < s->C || x->num_rows <= 0) return FVAE_ERR_SHAPE;
        return FVAE_OK;
    }
    if (x->row_pitch < s->C || x->seq_pitch < int64_t(s->T - 1) * x->row_pitch + s->C) return FVAE_ERR_SHAPE;
    return FVAE_OK;
}

int check_noise(const fvae_noise* nz, uint32_t flags) {
    if (!nz) return FVAE_ERR_NULL;
    if (!(flags & FVAE_FLAG_PHILOX)) {
        if (!nz->eps) return FVAE_ERR_NULL;
        if ((flags & FVAE_FLAG_TRAIN) && !nz->keep_mask) return FVAE_ERR_NULL;
    }
    return FVAE_OK;
}

void bind_params(const float* p, const Layout& L, FeW& fw, HeadsW& hw) {
    fw.ln_w = p + L.off[FVAE_P_LN_W];   fw.ln_b = p + L.off[FVAE_P_LN_B];
    fw.W1 = p + L.off[FVAE_P_W1];       fw.b1 = p + L.off[FVAE_P_B1];
    fw.Wih = p + L.off[FVAE_P_WIH];     fw.Whh = p + L.off[FVAE_P_WHH];
    fw.bih = p + L.off[FVAE_P_BIH];     fw.bhh = p + L.off[FVAE_P_BHH];
    hw.Wp = p + L.off[FVAE_P_ENC_W];    hw.bp = p + L.off[FVAE_P_ENC_B];
    hw.Wmu = p + L.off[FVAE_P_ENC_MU_W]; hw.bmu = p + L.off[FVAE_P_ENC_MU_B];
    hw.Wsig = p + L.off[FVAE_P_ENC_SG_W]; hw.bsig = p + L.off[FVAE_P_ENC_SG_B];
    hw.Wa = p + L.off[FVAE_P_AL_W];     hw.ba = p + L.off[FVAE_P_AL_B];
    hw.wam = p + L.off[FVAE_P_AL_MU_W]; hw.bam = p + L.off[FVAE_P_AL_MU_B];
    hw.was = p + L.off[FVAE_P_AL_SG_W]; hw.bas = p + L.off[FVAE_P_AL_SG_B];
    hw.Wb = p + L.off[FVAE_P_BETA_W];   hw.bb = p + L.off[FVAE_P_BETA_B];
    hw.q = p + L.off[FVAE_P_ATT_Q];
    hw.Wk = p + L.off[FVAE_P_ATT_KW];   hw.bk = p + L.off[FVAE_P_ATT_KB];
    hw.Wv = p + L.off[FVAE_P_ATT_VW];   hw.bv = p + L.off[FVAE_P_ATT_VB];
    hw.Wl = p + L.off[FVAE_P_PR_W];     hw.bl = p + L.off[FVAE_P_PR_B];
    hw.wpm = p + L.off[FVAE_P_PR_MU_W]; hw.bpm = p + L.off[FVAE_P_PR_MU_B];
    hw.wps = p + L.off[FVAE_P_PR_SG_W]; hw.bps = p + L.off[FVAE_P_PR_SG_B];
}

void bind_grads(float* p, const Layout& L, FeG& fg, HeadsG& hg) {
    FeW fw; HeadsW hw;
    bind_params(p, L, fw, hw);
    // same layout: reuse the const binding and cast the constness away for the gradient views
    fg = FeG{const_cast<float*>(fw.ln_w), const_cast<float*>(fw.ln_b), const_cast<float*>(fw.W1), const_cast<float*>(fw.b1),
             const_cast<float*>(fw.Wih), const_cast<float*>(fw.Whh), const_cast<float*>(fw.bih), const_cast<float*>(fw.bhh)};
    hg = HeadsG{const_cast<float*>(hw.Wp), const_cast<float*>(hw.bp), const_cast<float*>(hw.Wmu), const_cast<float*>(hw.bmu),
                const_cast<float*>(hw.Wsig), const_cast<float*>(hw.bsig), const_cast<float*>(hw.Wa), const_cast<float*>(hw.ba),
                const_cast<float*>(hw.wam), const_cast<float*>(hw.bam), const_cast<float*>(hw.was), const_cast<float*>(hw.bas),
                const_cast<float*>(hw.Wb), const_cast<float*>(hw.bb), const_cast<float*>(hw.q), const_cast<float*>(hw.Wk),
                const_cast<float*>(hw.bk), const_cast<float*>(hw.Wv), const_cast<float*>(hw.bv), const_cast<float*>(hw.Wl),
                const_cast<float*>(hw.bl), const_cast<float*>(hw.wpm), const_cast<float*>(hw.bpm), const_cast<float*>(hw.wps),
                const_cast<float*>(hw.bps)};
}

int check_outputs(const fvae_outputs* o, bool predict) {
    if (!o) return FVAE_ERR_NULL;
    if (!o->yhat || !o->mu_y || !o->sigma_y || !o->mu_prior || !o->sigma_prior) return FVAE_ERR_NULL;
    if (!predict && (!o->loss || !o->date_loss || !o->mu_post || !o->sigma_post)) return FVAE_ERR_NULL;
    return FVAE_OK;
}

int fe_forward_any(const FeDims& fd, const fvae_panel& x, const FeW& fw, int precision, float* e, void* ws, cudaStream_t st) {
    return precision == FVAE_PREC_BF16_TC ? fe_tc_forward(fd, x, fw, e, ws, st) : fe_f32_forward(fd, x, fw, e, ws, st);
}
int fe_backward_any(const FeDims& fd, const fvae_panel& x, const FeW& fw, const FeG& fg, int precision, const float* dE,
                    void* ws, cudaStream_t st) {
    return precision == FVAE_PREC_BF16_TC ? fe_tc_backward(fd, x, fw, fg, dE, ws, st)
                                          : fe_f32_backward(fd, x, fw, fg, dE, ws, st);
}

HeadsArgs make_heads_args(const fvae_shape& s, const Workspace& W, const float* y, const int32_t* date_ptr,
                          const fvae_noise& nz, uint32_t flags, int predict, const fvae_outputs& out, const HeadsW& hw) {
    HeadsArgs a;
    a.S = s.S; a.B = s.B; a.H = s.H; a.K = s.K; a.M = s.M;
    a.date_ptr = date_ptr; a.e = W.e; a.y = y; a.noise = nz; a.flags = flags; a.predict = predict;
    a.out = out; a.w = hw; a.sv = W.sv; a.use_tc = 0;
    a.parts = HeadsParts{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    return a;
}

int run_forward(const fvae_shape* shape, const fvae_panel* x, const float* y, const int32_t* date_ptr, const float* params,
                const fvae_noise* noise, uint32_t flags, int32_t precision, const fvae_outputs* out, void* workspace,
                int64_t workspace_bytes, void* stream, int predict) {
    int rc;
    if ((rc = check_shape(shape, precision)) != 0) return rc;
    if ((rc = check_panel(x, shape)) != 0) return rc;
    if ((rc = check_noise(noise, predict ? (flags & ~FVAE_FLAG_TRAIN) : flags)) != 0) return rc;
    if ((rc = check_outputs(out, predict != 0)) != 0) return rc;
    if (!date_ptr || !params || !workspace || (!predict && !y)) return FVAE_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return FVAE_ERR_WORKSPACE;
    Workspace W = carve(*shape, precision, workspace);
    if (W.bytes > workspace_bytes) return FVAE_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Layout L = make_layout(shape->C, shape->H, shape->K, shape->M);
    FeW fw; HeadsW hw;
    bind_params(params, L, fw, hw);
    if (predict) flags &= ~FVAE_FLAG_TRAIN;          // prediction runs the modules in eval() (utils.py:76)
    HeadsArgs a = make_heads_args(*shape, W, y, date_ptr, *noise, flags, predict, *out, hw);
    const FeDims fd{shape->S, shape->T, shape->C, shape->H};
    a.use_tc = (precision == FVAE_PREC_BF16_TC) ? 1 : 0;
    if ((rc = heads_prep(a, false, st)) != 0) return rc;
    if (a.use_tc && heads_tc_supported(shape->H, shape->K, shape->M) && (rc = heads_tc_prep(a, st)) != 0) return rc;
    if ((rc = fe_forward_any(fd, *x, fw, precision, W.e, W.fe_ws, st)) != 0) return rc;
    if ((rc = heads_forward(a, st)) != 0) return rc;
    if (!predict && (rc = loss_reduce(out->date_loss, shape->B, out->loss, st)) != 0) return rc;
    return FVAE_OK;
}

}  // namespace

extern "C" {

int fvae_abi_version(void) { return FVAE_ABI_VERSION; }

uint64_t fvae_debug_launch_count(void) { return __atomic_load_n(&fvae::g_launch_count, __ATOMIC_RELAXED); }

const char* fvae_status_string(int status) {
    switch (status) {
        case FVAE_OK: return "ok";
        case FVAE_ERR_NULL: return "a required pointer is NULL";
        case FVAE_ERR_SHAPE: return "invalid shape or pitch";
        case FVAE_ERR_LIMIT: return "outside the supported range (C<=192, H<=64, K<=256, M<=1024, shared memory)";
        case FVAE_ERR_DTYPE: return "unknown dtype or precision";
        case FVAE_ERR_WORKSPACE: return "workspace too small or not 256-byte aligned";
        case FVAE_ERR_NO_DEVICE: return "no CUDA device of the required architecture (sm_100a)";
        case FVAE_ERR_UNSUPPORTED: return "combination not implemented on this path";
        default: break;
    }
    if (status > 0) return cudaGetErrorString(static_cast<cudaError_t>(status));
    return "unknown status";
}

int fvae_param_offsets(int32_t C, int32_t H, int32_t K, int32_t M, int64_t* host_offsets) {
    if (!host_offsets) return FVAE_ERR_NULL;
    if (C <= 0 || H <= 0 || K <= 0 || M <= 0) return FVAE_ERR_SHAPE;
    const Layout L = make_layout(C, H, K, M);
    std::memcpy(host_offsets, L.off, sizeof(L.off));
    return FVAE_OK;
}

int64_t fvae_param_count(int32_t C, int32_t H, int32_t K, int32_t M) {
    if (C <= 0 || H <= 0 || K <= 0 || M <= 0) return FVAE_ERR_SHAPE;
    return make_layout(C, H, K, M).off[FVAE_P_NUM_SECTIONS];
}

int64_t fvae_workspace_bytes(const fvae_shape* shape, int32_t precision) {
    const int rc = check_shape(shape, precision);
    if (rc != 0) return rc;
    return carve(*shape, precision, nullptr).bytes;
}

int fvae_elbo_forward(const fvae_shape* shape, const fvae_panel* x, const float* y, const int32_t* date_ptr,
                      const float* params, const fvae_noise* noise, uint32_t flags, int32_t precision,
                      const fvae_outputs* out, void* workspace, int64_t workspace_bytes, void* stream) {
    return run_forward(shape, x, y, date_ptr, params, noise, flags, precision, out, workspace, workspace_bytes, stream, 0);
}

int fvae_predict(const fvae_shape* shape, const fvae_panel* x, const int32_t* date_ptr, const float* params,
                 const fvae_noise* noise, uint32_t flags, int32_t precision, const fvae_outputs* out, void* workspace,
                 int64_t workspace_bytes, void* stream) {
    return run_forward(shape, x, nullptr, date_ptr, params, noise, flags, precision, out, workspace, workspace_bytes, stream, 1);
}

int fvae_elbo_backward(const fvae_shape* shape, const fvae_panel* x, const float* y, const int32_t* date_ptr,
                       const float* params, const fvae_noise* noise, uint32_t flags, int32_t precision,
                       const fvae_outputs* out, float* grad, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc;
    if ((rc = check_shape(shape, precision)) != 0) return rc;
    if ((rc = check_panel(x, shape)) != 0) return rc;
    if ((rc = check_noise(noise, flags)) != 0) return rc;
    if ((rc = check_outputs(out, false)) != 0) return rc;
    if (!date_ptr || !params || !workspace || !y || !grad) return FVAE_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return FVAE_ERR_WORKSPACE;
    Workspace W = carve(*shape, precision, workspace);
    if (W.bytes > workspace_bytes) return FVAE_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Layout L = make_layout(shape->C, shape->H, shape->K, shape->M);
    FeW fw; HeadsW hw; FeG fg; HeadsG hg;
    bind_params(params, L, fw, hw);
    bind_grads(grad, L, fg, hg);
    HeadsArgs a = make_heads_args(*shape, W, y, date_ptr, *noise, flags, 0, *out, hw);
    a.use_tc = (precision == FVAE_PREC_BF16_TC) ? 1 : 0;
    const FeDims fd{shape->S, shape->T, shape->C, shape->H};
    cudaError_t ce = cudaMemsetAsync(grad, 0, size_t(L.off[FVAE_P_NUM_SECTIONS]) * sizeof(float), st);
    if (ce != cudaSuccess) return int(ce);
    // dG and dc are adjacent in the workspace (carve): one memset node
    if ((ce = cudaMemsetAsync(W.sv.dG, 0, size_t(reinterpret_cast<char*>(W.sv.dc) - reinterpret_cast<char*>(W.sv.dG)) + size_t(shape->K) * sizeof(float), st)) != cudaSuccess) return int(ce);
    if ((rc = heads_backward(a, hg, W.dE, st)) != 0) return rc;
    if ((rc = heads_post(a, hg, st)) != 0) return rc;
    if ((rc = fe_backward_any(fd, *x, fw, fg, precision, W.dE, W.fe_ws, st)) != 0) return rc;
    return FVAE_OK;
}

int fvae_fe_forward(const fvae_shape* shape, const fvae_panel* x, const float* params, int32_t precision, float* e,
                    void* workspace, int64_t workspace_bytes, void* stream) {
    int rc;
    if ((rc = check_shape(shape, precision)) != 0) return rc;
    if ((rc = check_panel(x, shape)) != 0) return rc;
    if (!params || !workspace || !e) return FVAE_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return FVAE_ERR_WORKSPACE;
    Workspace W = carve(*shape, precision, workspace);
    if (W.bytes > workspace_bytes) return FVAE_ERR_WORKSPACE;
    const Layout L = make_layout(shape->C, shape->H, shape->K, shape->M);
    FeW fw; HeadsW hw;
    bind_params(params, L, fw, hw);
    const FeDims fd{shape->S, shape->T, shape->C, shape->H};
    return fe_forward_any(fd, *x, fw, precision, e, W.fe_ws, static_cast<cudaStream_t>(stream));
}

int fvae_fe_backward(const fvae_shape* shape, const fvae_panel* x, const float* params, int32_t precision, const float* de,
                     float* grad, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc;
    if ((rc = check_shape(shape, precision)) != 0) return rc;
    if ((rc = check_panel(x, shape)) != 0) return rc;
    if (!params || !workspace || !de || !grad) return FVAE_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return FVAE_ERR_WORKSPACE;
    Workspace W = carve(*shape, precision, workspace);
    if (W.bytes > workspace_bytes) return FVAE_ERR_WORKSPACE;
    const Layout L = make_layout(shape->C, shape->H, shape->K, shape->M);
    FeW fw; HeadsW hw; FeG fg; HeadsG hg;
    bind_params(params, L, fw, hw);
    bind_grads(grad, L, fg, hg);
    const FeDims fd{shape->S, shape->T, shape->C, shape->H};
    return fe_backward_any(fd, *x, fw, fg, precision, de, W.fe_ws, static_cast<cudaStream_t>(stream));
}

int fvae_heads_parts(const fvae_shape* shape, const float* latent, const float* y, const int32_t* date_ptr, const float* params,
                     const fvae_noise* noise, uint32_t flags, const fvae_parts* parts, const fvae_outputs* out, void* workspace,
                     int64_t workspace_bytes, void* stream) {
    int rc;
    const int predict = y ? 0 : 1;
    if ((rc = check_shape(shape, FVAE_PREC_FP32)) != 0) return rc;
    if ((rc = check_noise(noise, flags)) != 0) return rc;
    if ((rc = check_outputs(out, predict != 0)) != 0) return rc;
    if (!latent || !date_ptr || !params || !workspace) return FVAE_ERR_NULL;
    if (parts && ((parts->z_mu == nullptr) != (parts->z_sigma == nullptr) || (parts->alpha_mu == nullptr) != (parts->alpha_sigma == nullptr)))
        return FVAE_ERR_NULL;
    if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0) return FVAE_ERR_WORKSPACE;
    Workspace W = carve(*shape, FVAE_PREC_FP32, workspace);
    if (W.bytes > workspace_bytes) return FVAE_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Layout L = make_layout(shape->C, shape->H, shape->K, shape->M);
    FeW fw; HeadsW hw;
    bind_params(params, L, fw, hw);
    HeadsArgs a = make_heads_args(*shape, W, y, date_ptr, *noise, flags, predict, *out, hw);
    a.e = latent;                                     // the caller's stock latents, read in place
    if (parts) a.parts = HeadsParts{parts->z_mu, parts->z_sigma, parts->alpha_mu, parts->alpha_sigma, parts->beta, parts->context};
    if ((rc = heads_prep(a, false, st)) != 0) return rc;
    if ((rc = heads_forward(a, st)) != 0) return rc;
    if (!predict && (rc = loss_reduce(out->date_loss, shape->B, out->loss, st)) != 0) return rc;
    return FVAE_OK;
}

int fvae_debug_front_forward(const fvae_shape* shape, const fvae_panel* x, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc;
    if ((rc = check_shape(shape, FVAE_PREC_BF16_TC)) != 0) return rc;
    if ((rc = check_panel(x, shape)) != 0) return rc;
    if (!workspace) return FVAE_ERR_NULL;
    Workspace W = carve(*shape, FVAE_PREC_BF16_TC, workspace);
    if (W.bytes > workspace_bytes) return FVAE_ERR_WORKSPACE;
    const FeDims fd{shape->S, shape->T, shape->C, shape->H};
    return fe_tc_front_only(fd, *x, W.fe_ws, static_cast<cudaStream_t>(stream));
}

const float* fvae_workspace_latent(const fvae_shape* shape, int32_t precision, const void* workspace) {
    if (check_shape(shape, precision) != 0 || !workspace) return nullptr;
    return carve(*shape, precision, const_cast<void*>(workspace)).e;
}

}  // extern "C"
