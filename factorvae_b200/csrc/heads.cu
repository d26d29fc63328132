// Cross-sectional heads of the FactorVAE ELBO step, fp32 CUDA-core kernels (sm_100a).
//
// One CTA per DATE.  Everything that couples the stocks of a date lives here:
//   FactorEncoder   (reference module.py:52-67, :44-50)  softmax over STOCKS, portfolio returns
//   FactorPredictor (module.py:169-188, AttentionLayer :134-153) -- the K heads are collapsed:
//         score_ik = (e_i . G_k + c_k)/sqrt(H+1e-6),  G_k = Wk_k^T q_k,  c_k = q_k . bk_k
//         ctx_k    = Wv_k (sum_i a_ik e_i) + bv_k
//   FactorDecoder   (module.py:107-123) alpha/beta heads, mu_y / sigma_y, reparameterisation
//   loss            (module.py:261-268) mean-squared error of the sample + KL(post || prior)
//
// Work decomposition inside a CTA: "thread per weight row".  All per-stock contractions have the
// form F[i][c] = e_i . w_c with w_c a length-H weight row (rows of Wp, G, Wb, Wa).  A thread owns
// one row c, keeps it in registers and sweeps the stocks of the date in chunks of 64 staged in
// shared memory (all lanes read the same e_i -> broadcast LDS.128).  Column softmaxes over stocks
// are online (running max / sum), so one sweep gives the encoder and attention statistics.
// Backward uses the same ownership for the weight gradients (dW_c = sum_i Z[i][c] e_i held in
// registers, flushed once per CTA with red.global.add) and a second mapping (thread per stock x
// quarter of H) for dE = Z . Wcat.
//
// HBM traffic: e (S x H fp32) is read twice forward / three times backward, from L2 in practice.
#include <float.h>
#include <stdlib.h>

#include <cooperative_groups.h>

#include "heads.cuh"

namespace fvae {

namespace {

constexpr int CH = 64;    // stocks per shared-memory chunk
constexpr int NT = 256;   // threads per CTA

__device__ __forceinline__ bool is_finite(float v) { return fabsf(v) <= FLT_MAX; }

template <int HP>
__device__ __forceinline__ float dot_row(const float (&w)[HP], const float* __restrict__ erow) {
    const float4* e4 = reinterpret_cast<const float4*>(erow);
    float acc = 0.f;
#pragma unroll
    for (int h4 = 0; h4 < HP / 4; ++h4) {
        float4 v = e4[h4];
        acc = fmaf(w[4 * h4 + 0], v.x, acc);
        acc = fmaf(w[4 * h4 + 1], v.y, acc);
        acc = fmaf(w[4 * h4 + 2], v.z, acc);
        acc = fmaf(w[4 * h4 + 3], v.w, acc);
    }
    return acc;
}

template <int HP>
__device__ __forceinline__ void load_row(float (&w)[HP], const float* __restrict__ src, int H, bool active) {
#pragma unroll
    for (int h = 0; h < HP; ++h) w[h] = (active && h < H) ? src[h] : 0.f;
}

struct Smem {
    float *Es, *ys, *aux0, *aux1, *red, *yp, *muz, *sgz, *mupr, *sgpr, *pooled, *ctx, *F;
    int* bad;
    int FLD;
};

// ------------------------------------------------------------------------------------------
// prep: collapse the K attention heads' key projections (date independent)
// ------------------------------------------------------------------------------------------
__global__ void heads_prep_kernel(HeadsArgs a, int zero_acc) {
    __shared__ float sWk[kMaxH * kMaxH];     // one coalesced pass over W_k, then the products run out of shared memory
    __shared__ float sq[kMaxH], sbk[kMaxH];
    const int k = blockIdx.x, H = a.H;
    const float* Wk = a.w.Wk + size_t(k) * H * H;
    for (int i = threadIdx.x; i < H * H; i += blockDim.x) sWk[i] = Wk[i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) { sq[i] = a.w.q[size_t(k) * H + i]; sbk[i] = a.w.bk[size_t(k) * H + i]; }
    __syncthreads();
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        float g = 0.f;
        for (int hp = 0; hp < H; ++hp) g = fmaf(sq[hp], sWk[hp * H + h], g);
        a.sv.G[k * H + h] = g;
        if (zero_acc) a.sv.dG[k * H + h] = 0.f;
    }
    if (threadIdx.x == blockDim.x - 1) {
        float c = 0.f;
        for (int h = 0; h < H; ++h) c = fmaf(sq[h], sbk[h], c);
        a.sv.cvec[k] = c;
        if (zero_acc) a.sv.dc[k] = 0.f;
    }
}

// backward of prep: dq, dWk, dbk from the accumulated dG, dc
__global__ void heads_post_kernel(HeadsArgs a, HeadsG g) {
    const int k = blockIdx.x, H = a.H;
    const float* q = a.w.q + size_t(k) * H;
    const float* Wk = a.w.Wk + size_t(k) * H * H;
    const float* bk = a.w.bk + size_t(k) * H;
    const float* dG = a.sv.dG + size_t(k) * H;
    const float dc = a.sv.dc[k];
    // a head that tripped the guard on every date has dG == dc == 0 exactly and must get exact
    // zeros even if q / Wk hold inf (0*inf): branch instead of multiplying.
    for (int idx = threadIdx.x; idx < H * H; idx += blockDim.x) {
        int hp = idx / H, h = idx % H;
        float d = dG[h];
        g.Wk[size_t(k) * H * H + idx] = (d == 0.f) ? 0.f : q[hp] * d;
    }
    for (int hp = threadIdx.x; hp < H; hp += blockDim.x) {
        float acc = 0.f;
        for (int h = 0; h < H; ++h) { float d = dG[h]; if (d != 0.f) acc = fmaf(Wk[hp * H + h], d, acc); }
        if (dc != 0.f) acc = fmaf(bk[hp], dc, acc);
        g.q[size_t(k) * H + hp] = acc;
        g.bk[size_t(k) * H + hp] = (dc == 0.f) ? 0.f : q[hp] * dc;
    }
}

__global__ void loss_reduce_kernel(const float* __restrict__ date_loss, int B, float* __restrict__ loss) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int d = threadIdx.x; d < B; d += blockDim.x) acc += date_loss[d];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) *loss = acc / float(B);
}

__device__ __forceinline__ Smem carve(float* smem, int HP, int H, int K, int M, int fld) {
    Smem s;
    s.Es = smem;                 smem += CH * HP;
    s.ys = smem;                 smem += CH;
    s.aux0 = smem;               smem += CH;
    s.aux1 = smem;               smem += CH;
    s.red = smem;                smem += 32;
    s.yp = smem;                 smem += M;
    s.muz = smem;                smem += K;
    s.sgz = smem;                smem += K;
    s.mupr = smem;               smem += K;
    s.sgpr = smem;               smem += K;
    s.bad = reinterpret_cast<int*>(smem); smem += K;
    s.pooled = smem;             smem += K * H;
    s.ctx = smem;                smem += K * H;
    s.F = smem;
    s.FLD = fld;
    return s;
}

template <int HP>
__device__ __forceinline__ void stage_chunk(const HeadsArgs& a, const Smem& s, int p0, int i0, int cn) {
    const int H = a.H;
    for (int idx = threadIdx.x; idx < CH * HP; idx += NT) {
        int i = idx / HP, h = idx % HP;
        s.Es[idx] = (i < cn && h < H) ? a.e[size_t(p0 + i0 + i) * H + h] : 0.f;
    }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
// Sweeps use threads as (weight row, stock slice) pairs: with ncol rows and NT threads there are
// NS = NT / ncol slices, slice s takes stocks s, s+NS, ... of each chunk, and the partial online-softmax
// states of a row are merged once per date.  Nobody idles at the chunk barriers.

__device__ __forceinline__ void merge_state(float& m, float& l, float m2, float l2, float& sc1, float& sc2) {
    const float mm = fmaxf(m, m2);
    sc1 = (m == mm) ? 1.f : expf(m - mm);      // also right when both are -inf / NaN-free by construction
    sc2 = (m2 == mm) ? 1.f : expf(m2 - mm);
    l = l * sc1 + l2 * sc2;
    m = mm;
}

// Under-filled grids (fewer dates than SMs: the micro-batches of BASELINE configs[3..4]) launch a thread-block CLUSTER per date,
// gridDim.y = cluster size CS: CTA `crank` sweeps the 64-stock chunks crank, crank + CS, ... in both passes, the per-column softmax
// states and the per-date sums are exchanged through distributed shared memory and merged in rank order by every CTA (identical
// bits everywhere), the small per-date phases run redundantly, rank 0 writes the per-date results.
template <int HP>
__global__ void __launch_bounds__(NT) heads_fwd_kernel(HeadsArgs a) {
    extern __shared__ __align__(16) float smem_raw[];
    namespace cg = cooperative_groups;
    const uint64_t nstep = noise_step(a);
    const int H = a.H, K = a.K, M = a.M;
    const int d = blockIdx.x, tid = threadIdx.x;
    const int CS = int(gridDim.y), crank = int(blockIdx.y);
    const bool lead = crank == 0;
    cg::cluster_group cluster = cg::this_cluster();
    const int p0 = a.date_ptr[d], n = a.date_ptr[d + 1] - p0;
    const int FLD = (K + H) | 1;
    Smem s = carve(smem_raw, HP, H, K, M, FLD);
    if (n <= 0) { if (tid == 0 && lead && a.out.date_loss) a.out.date_loss[d] = nanf(""); return; }
    const float tau = sqrtf(float(H) + 1e-6f);          // module.py:142, evaluated in fp32

    // ---- pass A: online column softmax over the stocks for encoder (M) and attention (K) columns
    const int colA0 = a.predict ? M : 0;
    const int ncolA = a.predict ? K : M + K;
    // Thread -> (column, stock slice).  Encoder columns are cheap (dot + exp) and get one thread each; attention
    // columns (mask hash, ReLU, exp, pooled e-accumulation) cost ~3x as much per stock and get the remaining
    // threads as NS slices each, so nobody idles at the chunk barriers.
    const int nenc = a.predict ? 0 : M;
    const bool one_batch = ncolA <= NT;
    const int cpb = one_batch ? ncolA : NT;                       // columns per batch
    int NS = 1;
    if (one_batch) { NS = (NT - nenc) / K; NS = NS < 1 ? 1 : (NS > 8 ? 8 : NS); }
    float* scrA = s.F;                                            // [K][NS][HP + 2]   (m, l, accp[HP])
    for (int cb = 0; cb < ncolA; cb += cpb) {
        int cl, slice;
        if (one_batch) {
            if (tid < nenc) { cl = tid; slice = 0; }
            else { cl = nenc + (tid - nenc) % K; slice = (tid - nenc) / K; }
        } else { cl = tid; slice = 0; }
        const int c = colA0 + cb + cl;
        const bool is_att = c >= M;
        const bool act = (cb + cl) < ncolA && slice < (is_att ? NS : 1);
        const int k = c - M;
        const int nsl = is_att ? NS : 1;
        float w[HP];
        load_row<HP>(w, is_att ? a.sv.G + size_t(act ? k : 0) * H : a.w.Wp + size_t(act ? c : 0) * H, H, act);
        const float bias = !act ? 0.f : (is_att ? a.sv.cvec[k] : a.w.bp[c]);
        float m = -INFINITY, l = 0.f, accy = 0.f;
        float accp[HP];
#pragma unroll
        for (int h = 0; h < HP; ++h) accp[h] = 0.f;
        int bad = 0;
        for (int i0 = crank * CH; i0 < n; i0 += CS * CH) {
            const int cn = min(CH, n - i0);
            __syncthreads();
            stage_chunk<HP>(a, s, p0, i0, cn);
            if (tid < CH) s.ys[tid] = (!a.predict && tid < cn) ? a.y[p0 + i0 + tid] : 0.f;
            __syncthreads();
            if (!act) continue;
            for (int i = slice; i < cn; i += nsl) {
                const float* er = s.Es + i * HP;
                float x = dot_row<HP>(w, er) + bias;
                if (is_att) {
                    x = x / tau;
                    x = x * keep_factor(a, nstep, p0 + i0 + i, k);
                    x = relu_nan(x);
                    if (!is_finite(x)) bad = 1;
                }
                float p;
                if (x > m) {                      // new running maximum: rescale what we have
                    const float sc = expf(m - x);
                    l *= sc; accy *= sc;
                    if (is_att) {
#pragma unroll
                        for (int h = 0; h < HP; ++h) accp[h] *= sc;
                    }
                    m = x; p = 1.f;
                } else {
                    p = expf(x - m);
                }
                l += p;
                if (is_att) {
#pragma unroll
                    for (int h = 0; h < HP; ++h) accp[h] = fmaf(p, er[h], accp[h]);
                } else {
                    accy = fmaf(p, s.ys[i], accy);
                }
            }
        }
        // ---- merge the NS partial states of each attention column
        __syncthreads();
        if (act && is_att && slice > 0) {
            float* o = scrA + (size_t(k) * NS + slice) * (HP + 2);
            o[0] = m; o[1] = l;
#pragma unroll
            for (int h = 0; h < HP; ++h) o[2 + h] = accp[h];
        }
        if (act && slice == 0 && is_att) s.bad[k] = 0;
        __syncthreads();
        if (act && bad) atomicOr(&s.bad[k], 1);
        __syncthreads();
        int bd = 0;
        if (act && slice == 0 && is_att) {
            for (int q = 1; q < NS; ++q) {
                float sc1, sc2;
                const float* o = scrA + (size_t(k) * NS + q) * (HP + 2);
                if (o[1] > 0.f || o[1] != o[1]) {          // that slice saw at least one stock
                    merge_state(m, l, o[0], o[1], sc1, sc2);
#pragma unroll
                    for (int h = 0; h < HP; ++h) accp[h] = accp[h] * sc1 + o[2 + h] * sc2;
                }
            }
            bd = s.bad[k];
        }
        if (CS > 1) {
            // cluster exchange of the per-column states: X[column of this batch][HP + 3] = {m, l, accy | guard, accp[HP]}
            constexpr int XS = HP + 3;
            __syncthreads();                                   // the slice scratch in F has been read
            float* X = s.F;
            if (act && slice == 0) {
                float* xr = X + size_t(cl) * XS;
                xr[0] = m; xr[1] = l; xr[2] = is_att ? float(bd) : accy;
                if (is_att) {
#pragma unroll
                    for (int h = 0; h < HP; ++h) xr[3 + h] = accp[h];
                }
            }
            cluster.sync();
            if (act && slice == 0) {
                m = -INFINITY; l = 0.f; accy = 0.f; bd = 0;
#pragma unroll
                for (int h = 0; h < HP; ++h) accp[h] = 0.f;
                for (int r = 0; r < CS; ++r) {                 // rank order: every CTA of the cluster gets identical bits
                    const float* o = cluster.map_shared_rank(X, r) + size_t(cl) * XS;
                    const float m2 = o[0], l2 = o[1], t2 = o[2];
                    if (is_att) bd |= (t2 != 0.f);
                    if (l2 > 0.f || l2 != l2) {
                        float sc1, sc2;
                        merge_state(m, l, m2, l2, sc1, sc2);
                        if (is_att) {
#pragma unroll
                            for (int h = 0; h < HP; ++h) accp[h] = accp[h] * sc1 + o[3 + h] * sc2;
                        } else {
                            accy = accy * sc1 + t2 * sc2;
                        }
                    }
                }
                if (is_att) s.bad[k] = bd;
            }
            cluster.sync();                                    // nobody reuses F while a peer still reads it
        }
        if (act && slice == 0) {
            if (is_att) {
                if (lead) {
                    a.sv.att_m[size_t(d) * K + k] = m;
                    a.sv.att_l[size_t(d) * K + k] = l;
                    a.sv.bad[size_t(d) * K + k] = bd;
                }
                const float inv = 1.f / l;
#pragma unroll
                for (int h = 0; h < HP; ++h)
                    if (h < H) {
                        const float pv = accp[h] * inv;
                        s.pooled[k * H + h] = pv;
                        if (lead) a.sv.pooled[(size_t(d) * K + k) * H + h] = pv;
                    }
            } else {
                const float v = accy / l;
                s.yp[c] = v;
                if (lead) {
                    a.sv.enc_m[size_t(d) * M + c] = m;
                    a.sv.enc_l[size_t(d) * M + c] = l;
                    a.sv.yp[size_t(d) * M + c] = v;
                }
            }
        }
    }
    __syncthreads();

    // ---- posterior (module.py:48-49) and the :117 clamp
    if (!a.predict) {
        for (int k = tid; k < K; k += NT) {
            float mu = a.w.bmu[k], pre = a.w.bsig[k];
            const float* wm = a.w.Wmu + size_t(k) * M;
            const float* ws = a.w.Wsig + size_t(k) * M;
            for (int j = 0; j < M; ++j) { mu = fmaf(wm[j], s.yp[j], mu); pre = fmaf(ws[j], s.yp[j], pre); }
            float sg = softplus(pre);
            const int cl = (sg == 0.f);
            if (cl) sg = kSigmaFloor;
            s.muz[k] = mu; s.sgz[k] = sg;
            if (lead) {
                a.out.mu_post[size_t(d) * K + k] = mu;
                a.out.sigma_post[size_t(d) * K + k] = sg;
                a.sv.pre_sg_post[size_t(d) * K + k] = pre;
                a.sv.clamp_post[size_t(d) * K + k] = cl;
            }
        }
    }
    // ---- prior: ctx_k = Wv_k pooled_k + bv_k (zeros if the guard tripped), shared MLP head
    for (int idx = tid; idx < K * H; idx += NT) {
        const int k = idx / H, j = idx % H;
        float v = 0.f;
        if (!s.bad[k]) {
            const float* wv = a.w.Wv + (size_t(k) * H + j) * H;
            v = a.w.bv[size_t(k) * H + j];
            for (int h = 0; h < H; ++h) v = fmaf(wv[h], s.pooled[k * H + h], v);
        }
        s.ctx[idx] = v;
        if (lead) {
            a.sv.ctx[size_t(d) * K * H + idx] = v;
            if (a.parts.context) a.parts.context[size_t(d) * K * H + idx] = v;
        }
    }
    __syncthreads();
    float* hm = s.F;     // [K][H] scratch (F is not live yet)
    for (int idx = tid; idx < K * H; idx += NT) {
        const int k = idx / H, j = idx % H;
        float v = a.w.bl[j];
        const float* wl = a.w.Wl + size_t(j) * H;
        for (int h = 0; h < H; ++h) v = fmaf(wl[h], s.ctx[k * H + h], v);
        if (lead) a.sv.hm_pre[size_t(d) * K * H + idx] = v;
        hm[idx] = lrelu(v);
    }
    __syncthreads();
    for (int k = tid; k < K; k += NT) {
        float mu = a.w.bpm[0], pre = a.w.bps[0];
        for (int j = 0; j < H; ++j) { mu = fmaf(a.w.wpm[j], hm[k * H + j], mu); pre = fmaf(a.w.wps[j], hm[k * H + j], pre); }
        float sg = softplus(pre);
        const int cl = (sg == 0.f);
        if (cl) sg = kSigmaFloor;                              // module.py:264-265 (and :117 in prediction)
        s.mupr[k] = mu; s.sgpr[k] = sg;
        if (lead) {
            a.out.mu_prior[size_t(d) * K + k] = mu;
            a.out.sigma_prior[size_t(d) * K + k] = sg;
            a.sv.pre_sg_prior[size_t(d) * K + k] = pre;
            a.sv.clamp_prior[size_t(d) * K + k] = cl;
        }
        if (a.predict) { s.muz[k] = mu; s.sgz[k] = sg; }
        if (a.parts.z_mu) {                                    // FactorDecoder.forward on its own: the caller's factors, :117 clamp
            const float zs = a.parts.z_sigma[size_t(d) * K + k];
            s.muz[k] = a.parts.z_mu[size_t(d) * K + k];
            s.sgz[k] = (zs == 0.f) ? kSigmaFloor : zs;
        }
    }
    __syncthreads();

    // ---- pass B: decoder per stock (module.py:109-123) + squared error of the sample
    float rec_part = 0.f;
    float c1a = 0.f, c1b = 0.f;                                    // sum_i beta_ik dmu_y_i, sum_i beta_ik^2 dvar_i (for backward)
    const float coefN = 2.f / (float(n) * float(a.B));
    const int ncolB = K + H;
    const int cpbB = ncolB < NT ? ncolB : NT;
    const int NSB = NT / cpbB;
    constexpr int PPS = NT / CH;                                   // threads per stock in the finishing step
    for (int i0 = crank * CH; i0 < n; i0 += CS * CH) {
        const int cn = min(CH, n - i0);
        __syncthreads();
        stage_chunk<HP>(a, s, p0, i0, cn);
        if (tid < CH) {
            s.ys[tid] = (!a.predict && tid < cn) ? a.y[p0 + i0 + tid] : 0.f;
            s.aux0[tid] = (tid < cn) ? eps_of(a, nstep, p0 + i0 + tid) : 0.f;
        }
        __syncthreads();
        for (int cb = 0; cb < ncolB; cb += cpbB) {
            const int cc = cb + tid % cpbB, slice = tid / cpbB;
            if (cc < ncolB && slice < NSB) {
                float w[HP];
                const bool is_beta = cc < K;
                load_row<HP>(w, is_beta ? a.w.Wb + size_t(cc) * H : a.w.Wa + size_t(cc - K) * H, H, true);
                const float bias = is_beta ? a.w.bb[cc] : a.w.ba[cc - K];
                for (int i = slice; i < cn; i += NSB) s.F[i * FLD + cc] = dot_row<HP>(w, s.Es + i * HP) + bias;
            }
        }
        __syncthreads();
        {
            const int i = tid / PPS, part = tid % PPS;
            float amu = 0.f, asp = 0.f, mu = 0.f, var = 0.f;
            if (i < cn) {
                const float* f = s.F + i * FLD;
                for (int j = part; j < H; j += PPS) {
                    const float ha = lrelu(f[K + j]);
                    amu = fmaf(a.w.wam[j], ha, amu);
                    asp = fmaf(a.w.was[j], ha, asp);
                }
                for (int k = part; k < K; k += PPS) {
                    const float b = f[k];
                    mu = fmaf(b, s.muz[k], mu);
                    var = fmaf(b * b, s.sgz[k] * s.sgz[k], var);
                }
            }
#pragma unroll
            for (int o = PPS / 2; o > 0; o >>= 1) {
                amu += __shfl_xor_sync(0xffffffffu, amu, o);
                asp += __shfl_xor_sync(0xffffffffu, asp, o);
                mu += __shfl_xor_sync(0xffffffffu, mu, o);
                var += __shfl_xor_sync(0xffffffffu, var, o);
            }
            if (i < cn && part == 0) {
                const float asig = softplus(asp + a.w.bas[0]);
                if (a.parts.alpha_mu) { a.parts.alpha_mu[p0 + i0 + i] = amu + a.w.bam[0]; a.parts.alpha_sigma[p0 + i0 + i] = asig; }
                mu += amu + a.w.bam[0];
                const float sy = sqrtf(var + asig * asig + 1e-6f);
                const float yh = fmaf(s.aux0[i], sy, mu);
                const int u = p0 + i0 + i;
                a.out.yhat[u] = yh; a.out.mu_y[u] = mu; a.out.sigma_y[u] = sy;
                const float dlt = yh - s.ys[i];
                rec_part = fmaf(dlt, dlt, rec_part);
                const float dmy = coefN * dlt;                     // d loss / d mu_y ; d loss / d sigma_y^2 = dmy eps / (2 sigma_y)
                s.aux1[i] = dmy;
                s.ys[i] = dmy * s.aux0[i] / (2.f * sy);
            } else if (i < CH && part == 0) {
                s.aux1[i] = 0.f; s.ys[i] = 0.f;
            }
        }
        if (a.parts.beta) {
            for (int idx = tid; idx < cn * K; idx += NT)
                a.parts.beta[size_t(p0 + i0 + idx / K) * K + idx % K] = s.F[(idx / K) * FLD + idx % K];
        }
        if (!a.predict) {
            __syncthreads();
            const int k = tid % K, slice = tid / K, NSC = NT / K;
            if (slice < NSC)
                for (int i = slice; i < cn; i += NSC) {
                    const float b = s.F[i * FLD + k];
                    c1a = fmaf(b, s.aux1[i], c1a);
                    c1b = fmaf(b * b, s.ys[i], c1b);
                }
        }
    }
    if (a.predict) return;
    __syncthreads();
    for (int i = tid; i < 2 * K; i += NT) s.F[i] = 0.f;
    __syncthreads();
    if (tid / K < NT / K) { atomicAdd(&s.F[tid % K], c1a); atomicAdd(&s.F[K + tid % K], c1b); }
    __syncthreads();
    float rec = block_sum(rec_part, s.red);
    if (CS > 1) {                                                       // per-date sums of the cluster, added in rank order by rank 0
        if (tid == 0) s.F[2 * K] = rec;
        cluster.sync();
        if (lead) {
            float rs = 0.f;
            for (int r = 0; r < CS; ++r) rs += cluster.map_shared_rank(s.F, r)[2 * K];
            rec = rs;
            for (int i = tid; i < 2 * K; i += NT) {
                float c1 = 0.f;
                for (int r = 0; r < CS; ++r) c1 += cluster.map_shared_rank(s.F, r)[i];
                if (i < K) a.sv.c1_mu[size_t(d) * K + i] = c1; else a.sv.c1_sg[size_t(d) * K + i - K] = c1;
            }
        }
        cluster.sync();                                                 // peers keep their shared memory until rank 0 has read it
        if (!lead) return;
    } else {
        for (int k = tid; k < K; k += NT) { a.sv.c1_mu[size_t(d) * K + k] = s.F[k]; a.sv.c1_sg[size_t(d) * K + k] = s.F[K + k]; }
    }
    rec /= float(n);                                                    // F.mse_loss: mean over stocks
    float klp = 0.f;
    for (int k = tid; k < K; k += NT) {                                 // module.py:247
        const float m1 = s.muz[k], s1 = s.sgz[k], m2 = s.mupr[k], s2 = s.sgpr[k];
        klp += logf(s2 / s1) + (s1 * s1 + (m1 - m2) * (m1 - m2)) / (2.f * s2 * s2) - 0.5f;
    }
    const float kl = block_sum(klp, s.red);
    if (tid == 0) a.out.date_loss[d] = rec + kl;
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
// Column space of the weight-row sweep: [0,M) encoder rows Wp | [M,M+K) collapsed attention rows G
// | [M+K,M+2K) beta rows Wb | [M+2K,M+2K+H) alpha rows Wa.
struct ColRef { const float* w; float bias; int kind; int idx; };
__device__ __forceinline__ ColRef col_ref(const HeadsArgs& a, int c) {
    const int H = a.H, K = a.K, M = a.M;
    ColRef r;
    if (c < M) { r.kind = 0; r.idx = c; r.w = a.w.Wp + size_t(c) * H; r.bias = a.w.bp[c]; }
    else if (c < M + K) { r.kind = 1; r.idx = c - M; r.w = a.sv.G + size_t(r.idx) * H; r.bias = a.sv.cvec[r.idx]; }
    else if (c < M + 2 * K) { r.kind = 2; r.idx = c - M - K; r.w = a.w.Wb + size_t(r.idx) * H; r.bias = a.w.bb[r.idx]; }
    else { r.kind = 3; r.idx = c - M - 2 * K; r.w = a.w.Wa + size_t(r.idx) * H; r.bias = a.w.ba[r.idx]; }
    return r;
}

// WSM: the stacked weight rows [Wp; G; Wb; Wa] are staged in shared memory for the dE sweep
// VEC: stop after the per-date vector phase and hand dyp / dp_k / pooled_k.dp_k to the tensor-core sweep (heads_tc.cu)
template <int HP, int NB, bool WSM, bool VEC = false>
__global__ void __launch_bounds__(NT) heads_bwd_kernel(HeadsArgs a, HeadsG g, float* __restrict__ dE) {
    const uint64_t nstep = noise_step(a);
    extern __shared__ __align__(16) float smem_raw[];
    const int H = a.H, K = a.K, M = a.M;
    const int d = blockIdx.x, tid = threadIdx.x;
    // blockIdx.y = slice of the date's stock chunks (the sweep over stocks needs only the per-date vectors: every slice
    // recomputes them, slice 0 alone adds the per-date parameter gradients).  One CTA per date left 84 of 148 SMs idle at
    // 64 dates per GPU (cfg4) -- the sweep was 52 % of that step.
    const int slice_y = blockIdx.y, nslice_y = gridDim.y;
    const bool lead = slice_y == 0;
    const int p0 = a.date_ptr[d], n = a.date_ptr[d + 1] - p0;
    if (n <= 0) return;
    if (slice_y * CH >= n) return;                          // nothing to sweep for this slice
    const int NF = M + 2 * K + H;
    const int ZLD = NF | 1;
    // shared memory
    float* sm = smem_raw;
    float* Es = sm;            sm += CH * HP;
    float* ys = sm;            sm += CH;
    float* dmy = sm;           sm += CH;     // d loss / d mu_y   per stock of the chunk
    float* dvv = sm;           sm += CH;     // d loss / d (sigma_y^2)
    float* damu = sm;          sm += CH;
    float* dasp = sm;          sm += CH;
    float* red = sm;           sm += 32;
    float* yp = sm;            sm += M;
    float* dyp = sm;           sm += M;
    float* encm = sm;          sm += M;
    float* encl = sm;          sm += M;
    float* muz = sm;           sm += K;
    float* sgz = sm;           sm += K;
    float* attm = sm;          sm += K;
    float* attl = sm;          sm += K;
    float* pdp = sm;           sm += K;      // pooled_k . dp_k
    float* dmupost = sm;       sm += K;
    float* dprepost = sm;      sm += K;
    float* dmuprior = sm;      sm += K;
    float* dpreprior = sm;     sm += K;
    int* bad = reinterpret_cast<int*>(sm); sm += K;
    float* pooled = sm;        sm += K * H;
    float* dps = sm;           sm += K * H;  // dp_k
    float* tmpKH = sm;         sm += K * H;  // dhm_pre, then dctx
    float* Aatt = sm;          sm += CH * (K | 1);
    float* Z = sm;             sm += CH * ZLD;
    float* Wc = sm;                           // WSM: [NF][HP] stacked weight rows (G rows pre-divided? no: plain)
    const int ALD = K | 1;

    const float tau = sqrtf(float(H) + 1e-6f);
    const float coefB = 1.f / float(a.B);
    const float coefN = 2.f / (float(n) * float(a.B));   // d/dyhat of mean-squared error, times 1/B

    // ---- per-date vectors
    for (int j = tid; j < M; j += NT) {
        yp[j] = a.sv.yp[size_t(d) * M + j];
        encm[j] = a.sv.enc_m[size_t(d) * M + j];
        encl[j] = a.sv.enc_l[size_t(d) * M + j];
    }
    for (int k = tid; k < K; k += NT) {
        muz[k] = a.out.mu_post[size_t(d) * K + k];
        sgz[k] = a.out.sigma_post[size_t(d) * K + k];
        attm[k] = a.sv.att_m[size_t(d) * K + k];
        attl[k] = a.sv.att_l[size_t(d) * K + k];
        bad[k] = a.sv.bad[size_t(d) * K + k];
        dmupost[k] = 0.f; dprepost[k] = 0.f;
    }
    for (int idx = tid; idx < K * H; idx += NT) pooled[idx] = a.sv.pooled[size_t(d) * K * H + idx];
    if (WSM) {
        for (int idx = tid; idx < NF * HP; idx += NT) {
            const int c = idx / HP, h = idx % HP;
            float v = 0.f;
            if (h < H) {
                if (c < M) v = a.w.Wp[size_t(c) * H + h];
                else if (c < M + K) v = a.sv.G[size_t(c - M) * H + h];
                else if (c < M + 2 * K) v = a.w.Wb[size_t(c - M - K) * H + h];
                else v = a.w.Wa[size_t(c - M - 2 * K) * H + h];
            }
            Wc[idx] = v;
        }
    }
    __syncthreads();

    // ---- d mu_z[k] = sum_i beta_ik dmu_y_i ; d sigma_z[k] = 2 sigma_z[k] sum_i beta_ik^2 dvar_i  (sums saved by forward pass B)
    for (int k = tid; k < K; k += NT) {
        dmupost[k] = a.sv.c1_mu[size_t(d) * K + k];
        dprepost[k] = 2.f * sgz[k] * a.sv.c1_sg[size_t(d) * K + k];
    }
    __syncthreads();

    // ---- vector phase: KL (module.py:247), mapping layer (:48-49), predictor head (:181-187)
    for (int k = tid; k < K; k += NT) {
        const float m1 = muz[k], s1 = sgz[k];
        const float m2 = a.out.mu_prior[size_t(d) * K + k], s2 = a.out.sigma_prior[size_t(d) * K + k];
        const float dm = m1 - m2;
        const float g_m1 = coefB * dm / (s2 * s2);
        const float g_s1 = coefB * (-1.f / s1 + s1 / (s2 * s2));
        const float g_s2 = coefB * (1.f / s2 - (s1 * s1 + dm * dm) / (s2 * s2 * s2));
        const float dmu1 = dmupost[k] + g_m1;
        const float dsg1 = a.sv.clamp_post[size_t(d) * K + k] ? 0.f : (dprepost[k] + g_s1);
        const float dpre1 = dsg1 * softplus_grad(a.sv.pre_sg_post[size_t(d) * K + k]);
        const float dsg2 = a.sv.clamp_prior[size_t(d) * K + k] ? 0.f : g_s2;
        const float dpre2 = dsg2 * softplus_grad(a.sv.pre_sg_prior[size_t(d) * K + k]);
        dmupost[k] = dmu1; dprepost[k] = dpre1; dmuprior[k] = -g_m1; dpreprior[k] = dpre2;
        if (lead) { atomicAdd(g.bmu + k, dmu1); atomicAdd(g.bsig + k, dpre1); }
    }
    __syncthreads();
    if (!lead) {
        // per-date parameter gradients belong to slice 0
    } else if ((M & 3) == 0) {                              // dWmu, dWsig: one vector reduction per 4 portfolios
        for (int i4 = tid; i4 < K * M / 4; i4 += NT) {
            const int idx = 4 * i4, k = idx / M, j = idx % M;
            const float a1 = dmupost[k], a2 = dprepost[k];
            red_add_v4(g.Wmu + idx, a1 * yp[j], a1 * yp[j + 1], a1 * yp[j + 2], a1 * yp[j + 3]);
            red_add_v4(g.Wsig + idx, a2 * yp[j], a2 * yp[j + 1], a2 * yp[j + 2], a2 * yp[j + 3]);
        }
    } else {
        for (int idx = tid; idx < K * M; idx += NT) {
            const int k = idx / M, j = idx % M;
            atomicAdd(g.Wmu + idx, dmupost[k] * yp[j]);
            atomicAdd(g.Wsig + idx, dprepost[k] * yp[j]);
        }
    }
    for (int j = tid; j < M; j += NT) {                    // d y_p
        float v = 0.f;
        for (int k = 0; k < K; ++k) {
            v = fmaf(a.w.Wmu[size_t(k) * M + j], dmupost[k], v);
            v = fmaf(a.w.Wsig[size_t(k) * M + j], dprepost[k], v);
        }
        dyp[j] = v;
    }
    // predictor head: hm = lrelu(hm_pre); mu_prior = wpm.hm + bpm; pre_prior = wps.hm + bps
    {
        float sbm = 0.f, sbs = 0.f;
        for (int k = tid; k < K; k += NT) { sbm += dmuprior[k]; sbs += dpreprior[k]; }
        sbm = block_sum(sbm, red);
        sbs = block_sum(sbs, red);
        if (tid == 0 && lead) { atomicAdd(g.bpm, sbm); atomicAdd(g.bps, sbs); }
    }
    // hm_pre and ctx of the date through shared memory: read from global inside the k loops below they were K dependent L2 round
    // trips per thread (37 % of this phase's stall samples)
    for (int idx = tid; idx < K * H; idx += NT) { tmpKH[idx] = a.sv.hm_pre[size_t(d) * K * H + idx]; dps[idx] = a.sv.ctx[size_t(d) * K * H + idx]; }
    __syncthreads();
    for (int j = tid; j < H; j += NT) {
        float a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float wm = a.w.wpm[j], ws = a.w.wps[j];
        for (int k = 0; k < K; ++k) {
            const float pre = tmpKH[k * H + j];
            const float hmv = lrelu(pre);
            a1 = fmaf(dmuprior[k], hmv, a1);
            a2 = fmaf(dpreprior[k], hmv, a2);
            const float dpre = (dmuprior[k] * wm + dpreprior[k] * ws) * (pre > 0.f ? 1.f : kLeakySlope);
            tmpKH[k * H + j] = dpre;                         // d hm_pre
            a3 += dpre;
        }
        if (lead) { atomicAdd(g.wpm + j, a1); atomicAdd(g.wps + j, a2); atomicAdd(g.bl + j, a3); }
    }
    __syncthreads();
    for (int idx = tid; lead && idx < H * H; idx += NT) {  // dWl[j][h] = sum_k dhm_pre[k][j] ctx[k][h]
        const int j = idx / H, h = idx % H;
        float v = 0.f;
        for (int k = 0; k < K; ++k) v = fmaf(tmpKH[k * H + j], dps[k * H + h], v);
        atomicAdd(g.Wl + idx, v);
    }
    __syncthreads();                                       // dps (ctx) has been read: it becomes scratch below
    // d ctx[k][h] = sum_j Wl[j][h] dhm_pre[k][j]   (zero behind a tripped guard) -> dps as scratch
    for (int idx = tid; idx < K * H; idx += NT) {
        const int k = idx / H, h = idx % H;
        float v = 0.f;
        if (!bad[k])
            for (int j = 0; j < H; ++j) v = fmaf(a.w.Wl[size_t(j) * H + h], tmpKH[k * H + j], v);
        dps[idx] = v;
    }
    __syncthreads();
    for (int idx = tid; idx < K * H; idx += NT) tmpKH[idx] = dps[idx];   // tmpKH := d ctx
    __syncthreads();
    for (int idx = tid; idx < K * H; idx += NT) {          // dbv, dp_k = Wv_k^T dctx_k
        const int k = idx / H, h = idx % H;
        float v = 0.f;
        if (!bad[k]) {
            if (lead) atomicAdd(g.bv + idx, tmpKH[idx]);
            const float* wv = a.w.Wv + size_t(k) * H * H;
            for (int j = 0; j < H; ++j) v = fmaf(wv[j * H + h], tmpKH[k * H + j], v);
        }
        dps[idx] = v;
    }
    if (!lead) {
    } else if ((H & 3) == 0) {                              // dWv[k][j][h] = dctx[k][j] pooled[k][h], 4 h per reduction
        for (int i4 = tid; i4 < K * H * H / 4; i4 += NT) {
            const int idx = 4 * i4, k = idx / (H * H), r = idx % (H * H), j = r / H, h = r % H;
            if (!bad[k]) {
                const float dc = tmpKH[k * H + j];
                const float* pl = pooled + k * H + h;
                red_add_v4(g.Wv + idx, dc * pl[0], dc * pl[1], dc * pl[2], dc * pl[3]);
            }
        }
    } else {
        for (int idx = tid; idx < K * H * H; idx += NT) {
            const int k = idx / (H * H), r = idx % (H * H), j = r / H, h = r % H;
            if (!bad[k]) atomicAdd(g.Wv + idx, tmpKH[k * H + j] * pooled[k * H + h]);
        }
    }
    __syncthreads();
    for (int k = tid; k < K; k += NT) {
        float v = 0.f;
        if (!bad[k]) for (int h = 0; h < H; ++h) v = fmaf(pooled[k * H + h], dps[k * H + h], v);
        pdp[k] = v;
    }
    __syncthreads();
    if (VEC) {
        for (int j = tid; j < M; j += NT) a.sv.t_dyp[size_t(d) * M + j] = dyp[j];
        for (int idx = tid; idx < K * H; idx += NT) a.sv.t_dps[size_t(d) * K * H + idx] = dps[idx];
        for (int k = tid; k < K; k += NT) a.sv.t_pdp[size_t(d) * K + k] = pdp[k];
        return;
    }

    // ---- pass C2: Z sweep -> weight gradients (registers) and dE
    // thread = (weight row c, stock slice): rows per batch cpb2, NS2 slices, NB batches of rows
    // NB == 1 (all rows fit one batch): encoder / beta / alpha rows get one thread, attention rows (two sweeps,
    // mask hash, exp) get NS2 slices each.  NB > 1: plain row = b*NT + tid.
    int NS2 = 1;
    if (NB == 1) { NS2 = (NT - M - K - H) / K; NS2 = NS2 < 1 ? 1 : (NS2 > 8 ? 8 : NS2); }
    int mycl = tid, myslice = 0;
    bool slice_ok = true;
    if (NB == 1) {
        if (tid < M) { mycl = tid; }
        else if (tid < M + K * NS2) { mycl = M + (tid - M) % K; myslice = (tid - M) / K; }
        else { mycl = M + K + (tid - M - K * NS2); slice_ok = mycl < NF; }
    }
    const int cpb2 = NT;
    const int mynsl = (NB == 1 && mycl >= M && mycl < M + K) ? NS2 : 1;
    const int cpbA = H < NT ? H : NT;                     // alpha rows in S1c
    const int NSA = NT / cpbA;
    constexpr int PPS = NT / CH;
    float acc[NB][HP];
    float accb[NB];
    float acc_wam = 0.f, acc_was = 0.f;       // S1c threads only
    float sum_damu = 0.f, sum_dasp = 0.f;     // per-stock threads
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        accb[b] = 0.f;
#pragma unroll
        for (int h = 0; h < HP; ++h) acc[b][h] = 0.f;
    }
    for (int i0 = slice_y * CH; i0 < n; i0 += nslice_y * CH) {
        const int cn = min(CH, n - i0);
        __syncthreads();
        stage_chunk<HP>(a, Smem{Es}, p0, i0, cn);
        if (tid < CH) {
            float v1 = 0.f, v2 = 0.f, yv = 0.f;
            if (tid < cn) {
                const int u = p0 + i0 + tid;
                yv = a.y[u];
                v1 = coefN * (a.out.yhat[u] - yv);
                v2 = v1 * eps_of(a, nstep, u) / (2.f * a.out.sigma_y[u]);
            }
            ys[tid] = yv; dmy[tid] = v1; dvv[tid] = v2;
        }
        __syncthreads();
        // S1: F = e . w_c (+bias) and its transform into the backward coefficient Z[i][c]
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int c = b * cpb2 + mycl;
            if (c < NF && slice_ok) {
                const ColRef cr = col_ref(a, c);
                float w[HP];
                if (cr.kind == 1 && bad[cr.idx]) {
                    for (int i = myslice; i < cn; i += mynsl) { Z[i * ZLD + c] = 0.f; Aatt[i * ALD + cr.idx] = 0.f; }
                } else {
                    load_row<HP>(w, cr.w, H, true);
                    if (cr.kind == 0) {                         // encoder: dlogit = w_ij dyp_j (y_i - yp_j)
                        const float mj = encm[c], lj = encl[c], dj = dyp[c], ypj = yp[c];
                        for (int i = myslice; i < cn; i += mynsl) {
                            const float x = dot_row<HP>(w, Es + i * HP) + cr.bias;
                            Z[i * ZLD + c] = expf(x - mj) / lj * dj * (ys[i] - ypj);
                        }
                    } else if (cr.kind == 1) {                  // attention: a_ik now, ds_ik after the dp sweep
                        const int k = cr.idx;
                        const float mk = attm[k], lk = attl[k];
                        for (int i = myslice; i < cn; i += mynsl) {
                            float x = (dot_row<HP>(w, Es + i * HP) + cr.bias) / tau;
                            const float kf = keep_factor(a, nstep, p0 + i0 + i, k);
                            x = x * kf;
                            const float r = relu_nan(x);
                            const float aik = expf(r - mk) / lk;
                            Aatt[i * ALD + k] = aik;
                            Z[i * ZLD + c] = (x > 0.f) ? kf / tau : 0.f;   // d relu(dropout(s)) / d (e.G+c)
                        }
                        load_row<HP>(w, dps + k * H, H, true);   // second sweep with dp_k
                        const float pk = pdp[k];
                        for (int i = myslice; i < cn; i += mynsl) {
                            const float dr = Aatt[i * ALD + k] * (dot_row<HP>(w, Es + i * HP) - pk);
                            Z[i * ZLD + c] *= dr;
                        }
                    } else if (cr.kind == 2) {                  // beta: mu_z dmu_y + 2 beta sigma_z^2 dv
                        const float mz = muz[cr.idx], sz2 = sgz[cr.idx] * sgz[cr.idx];
                        for (int i = myslice; i < cn; i += mynsl) {
                            const float bt = dot_row<HP>(w, Es + i * HP) + cr.bias;
                            Z[i * ZLD + c] = mz * dmy[i] + 2.f * bt * sz2 * dvv[i];
                        }
                    } else {                                    // alpha hidden: keep pre-activation for S1b/S1c
                        for (int i = myslice; i < cn; i += mynsl) Z[i * ZLD + c] = dot_row<HP>(w, Es + i * HP) + cr.bias;
                    }
                }
            }
        }
        __syncthreads();
        // S1b: per stock, alpha scalars: asig = softplus(was . lrelu(ha_pre) + bas); PPS threads per stock
        {
            const int i = tid / PPS, part = tid % PPS;
            float asp = 0.f;
            if (i < cn) {
                const float* zr = Z + i * ZLD + (M + 2 * K);
                for (int j = part; j < H; j += PPS) asp = fmaf(a.w.was[j], lrelu(zr[j]), asp);
            }
#pragma unroll
            for (int o = PPS / 2; o > 0; o >>= 1) asp += __shfl_xor_sync(0xffffffffu, asp, o);
            if (part == 0) {
                float v1 = 0.f, v2 = 0.f;
                if (i < cn) {
                    asp += a.w.bas[0];
                    const float asig = softplus(asp);
                    v1 = dmy[i];                                     // d alpha_mu
                    v2 = 2.f * asig * dvv[i] * softplus_grad(asp);   // d (pre-softplus alpha_sigma)
                    sum_damu += v1; sum_dasp += v2;
                }
                damu[i] = v1; dasp[i] = v2;
            }
        }
        __syncthreads();
        // S1c: alpha rows: mu/sigma layer weight grads, then Z := d ha_pre      thread = (alpha row j, slice)
        for (int jb = 0; jb < H; jb += cpbA) {
            const int j = jb + tid % cpbA, sl = tid / cpbA;
            if (j < H && sl < NSA) {
                const int c = M + 2 * K + j;
                const float wm = a.w.wam[j], ws = a.w.was[j];
                for (int i = sl; i < cn; i += NSA) {
                    const float pre = Z[i * ZLD + c];
                    const float ha = lrelu(pre);
                    acc_wam = fmaf(damu[i], ha, acc_wam);
                    acc_was = fmaf(dasp[i], ha, acc_was);
                    Z[i * ZLD + c] = (damu[i] * wm + dasp[i] * ws) * (pre > 0.f ? 1.f : kLeakySlope);
                }
            }
        }
        __syncthreads();
        // S2: weight-row gradients  dW_c += sum_i Z[i][c] e_i ; bias grads
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int c = b * cpb2 + mycl;
            if (c < NF && slice_ok) {
                for (int i = myslice; i < cn; i += mynsl) {
                    const float z = Z[i * ZLD + c];
                    accb[b] += z;
                    const float4* e4 = reinterpret_cast<const float4*>(Es + i * HP);
#pragma unroll
                    for (int h4 = 0; h4 < HP / 4; ++h4) {
                        const float4 v = e4[h4];
                        acc[b][4 * h4 + 0] = fmaf(z, v.x, acc[b][4 * h4 + 0]);
                        acc[b][4 * h4 + 1] = fmaf(z, v.y, acc[b][4 * h4 + 1]);
                        acc[b][4 * h4 + 2] = fmaf(z, v.z, acc[b][4 * h4 + 2]);
                        acc[b][4 * h4 + 3] = fmaf(z, v.w, acc[b][4 * h4 + 3]);
                    }
                }
            }
        }
        // S3: dE[i][:] = sum_c Z[i][c] Wcat[c][:] + sum_k a_ik dp_k    (thread: stock i, part of H)
        {
            constexpr int HQ = HP / PPS;
            const int i = tid % CH, hq = tid / CH;
            if (i < cn) {
                float o[HQ];
#pragma unroll
                for (int t = 0; t < HQ; ++t) o[t] = 0.f;
                const float* zr = Z + i * ZLD;
                for (int c = 0; c < NF; ++c) {
                    const float z = zr[c];
                    if (z != 0.f) {                       // also keeps 0*inf of a tripped head out
                        const float* wr;
                        if (WSM) wr = Wc + c * HP;
                        else if (c < M) wr = a.w.Wp + size_t(c) * H;
                        else if (c < M + K) wr = a.sv.G + size_t(c - M) * H;
                        else if (c < M + 2 * K) wr = a.w.Wb + size_t(c - M - K) * H;
                        else wr = a.w.Wa + size_t(c - M - 2 * K) * H;
#pragma unroll
                        for (int t = 0; t < HQ; ++t) {
                            const int h = hq * HQ + t;
                            if (WSM || h < H) o[t] = fmaf(z, wr[h], o[t]);
                        }
                    }
                }
                const float* ar = Aatt + i * ALD;
                for (int k = 0; k < K; ++k) {
                    const float av = ar[k];
#pragma unroll
                    for (int t = 0; t < HQ; ++t) {
                        const int h = hq * HQ + t;
                        if (h < H) o[t] = fmaf(av, dps[k * H + h], o[t]);
                    }
                }
#pragma unroll
                for (int t = 0; t < HQ; ++t) {
                    const int h = hq * HQ + t;
                    if (h < H) dE[size_t(p0 + i0 + i) * H + h] = o[t];
                }
            }
        }
    }
    // ---- flush the register accumulators (one red.global.add per parameter per (CTA, slice))
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int c = b * cpb2 + mycl;
        if (c < NF && slice_ok) {
            float* gw; float* gb;
            if (c < M) { gw = g.Wp + size_t(c) * H; gb = g.bp + c; }
            else if (c < M + K) { gw = a.sv.dG + size_t(c - M) * H; gb = a.sv.dc + (c - M); }
            else if (c < M + 2 * K) { gw = g.Wb + size_t(c - M - K) * H; gb = g.bb + (c - M - K); }
            else { gw = g.Wa + size_t(c - M - 2 * K) * H; gb = g.ba + (c - M - 2 * K); }
            const bool skip = (c >= M && c < M + K) && bad[c - M];
            if (!skip) {
#pragma unroll
                for (int h = 0; h < HP; ++h) if (h < H) atomicAdd(gw + h, acc[b][h]);
                atomicAdd(gb, accb[b]);
            }
        }
    }
    for (int jb = 0; jb < H; jb += cpbA) {
        const int j = jb + tid % cpbA, sl = tid / cpbA;
        if (j < H && sl < NSA && jb == 0) { atomicAdd(g.wam + j, acc_wam); atomicAdd(g.was + j, acc_was); }
    }
    sum_damu = block_sum(sum_damu, red);
    sum_dasp = block_sum(sum_dasp, red);
    if (tid == 0) { atomicAdd(g.bam, sum_damu); atomicAdd(g.bas, sum_dasp); }
}

size_t fwd_smem_bytes(int HP, int H, int K, int M) {
    size_t f = size_t(CH) * HP + 3 * CH + 32 + M + 5 * size_t(K) + 2 * size_t(K) * H;
    size_t fl = size_t(CH) * ((K + H) | 1);
    size_t hm = size_t(K) * H;
    size_t mg = size_t(NT) * (HP + 2);                                       // merge scratch of the statistics sweep
    size_t xs = size_t(M + K < NT ? M + K : NT) * (HP + 3);                  // cluster exchange of the column states
    size_t mx = fl > hm ? fl : hm;
    if (mg > mx) mx = mg;
    if (xs > mx) mx = xs;
    if (size_t(2 * K + 1) > mx) mx = size_t(2 * K + 1);
    return (f + mx) * sizeof(float);
}
size_t bwd_vec_smem_bytes(int HP, int H, int K, int M) {
    return (size_t(CH) * HP + 5 * CH + 32 + 4 * size_t(M) + 10 * size_t(K) + 3 * size_t(K) * H) * sizeof(float);
}
size_t bwd_smem_bytes(int HP, int H, int K, int M, bool wsm) {
    size_t f = size_t(CH) * HP + 5 * CH + 32 + 4 * size_t(M) + 10 * size_t(K) + 3 * size_t(K) * H
             + size_t(CH) * (K | 1) + size_t(CH) * ((M + 2 * K + H) | 1);
    if (wsm) f += size_t(M + 2 * K + H) * HP;
    return f * sizeof(float);
}

template <typename KernelT>
int set_smem(KernelT kernel, size_t bytes) {
    if (bytes > 227 * 1024) return FVAE_ERR_LIMIT;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    return int(e);
}

}  // namespace

int heads_prep(const HeadsArgs& a, bool zero_grad_acc, cudaStream_t stream) {
    heads_prep_kernel<<<a.K, 128, 0, stream>>>(a, zero_grad_acc ? 1 : 0); count_launch();
    return int(cudaGetLastError());
}

int heads_post(const HeadsArgs& a, const HeadsG& g, cudaStream_t stream) {
    heads_post_kernel<<<a.K, 128, 0, stream>>>(a, g); count_launch();
    return int(cudaGetLastError());
}

int loss_reduce(const float* date_loss, int B, float* loss, cudaStream_t stream) {
    loss_reduce_kernel<<<1, 256, 0, stream>>>(date_loss, B, loss); count_launch();
    return int(cudaGetLastError());
}

// padded hidden width: the sweeps cost HP FMAs per (stock, weight row), so pad H as little as possible
static int pick_hp(int H) { return H <= 20 ? 20 : (H <= 32 ? 32 : (H <= 48 ? 48 : 64)); }

int heads_forward(const HeadsArgs& a, cudaStream_t stream) {
    if (a.use_tc && heads_tc_supported(a.H, a.K, a.M)) return heads_tc_forward(a, stream);
    const int HP = pick_hp(a.H);
    const size_t smem = fwd_smem_bytes(HP, a.H, a.K, a.M);
    int rc;
    // fewer dates than SMs: a cluster of CS CTAs per date (power of two, at most 8 = the portable cluster size, at most one CTA
    // per 64-stock chunk of an average date, at most what fills the SMs once)
    int nsm = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int avg_chunks = (int((int64_t(a.S) + a.B - 1) / a.B) + CH - 1) / CH;
    int cs = 1;
    while (cs < 8 && 2 * cs * a.B <= nsm && 2 * cs <= avg_chunks) cs *= 2;
    if (getenv("FVAE_HEADS_CLUSTER")) { const int v = atoi(getenv("FVAE_HEADS_CLUSTER")); if (v == 1 || v == 2 || v == 4 || v == 8) cs = v; }
#define FVAE_LAUNCH_FWD(HPV)                                                        \
    do {                                                                            \
        if ((rc = set_smem(heads_fwd_kernel<HPV>, smem)) != 0) return rc;           \
        if (cs == 1) {                                                              \
            heads_fwd_kernel<HPV><<<a.B, NT, smem, stream>>>(a);                    \
        } else {                                                                    \
            cudaLaunchConfig_t cfg = {};                                            \
            cfg.gridDim = dim3(a.B, cs); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = stream; \
            cudaLaunchAttribute at[1];                                              \
            at[0].id = cudaLaunchAttributeClusterDimension;                         \
            at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = cs; at[0].val.clusterDim.z = 1; \
            cfg.attrs = at; cfg.numAttrs = 1;                                       \
            cudaError_t le = cudaLaunchKernelEx(&cfg, heads_fwd_kernel<HPV>, a);    \
            if (le != cudaSuccess) return int(le);                                  \
        }                                                                           \
        count_launch();                                                             \
    } while (0)
    if (HP == 20) FVAE_LAUNCH_FWD(20); else if (HP == 32) FVAE_LAUNCH_FWD(32); else if (HP == 48) FVAE_LAUNCH_FWD(48); else FVAE_LAUNCH_FWD(64);
#undef FVAE_LAUNCH_FWD
    return int(cudaGetLastError());
}

int heads_backward(const HeadsArgs& a, const HeadsG& g, float* dE, cudaStream_t stream) {
    const int HP = pick_hp(a.H);
    if (a.use_tc && heads_tc_supported(a.H, a.K, a.M)) {
        // vector phase per date on CUDA cores, then the stock sweep on tcgen05
        const size_t vsm = bwd_vec_smem_bytes(HP, a.H, a.K, a.M);
        int rcv;
        if (HP == 20) {
            if ((rcv = set_smem(heads_bwd_kernel<20, 1, false, true>, vsm)) != 0) return rcv;
            heads_bwd_kernel<20, 1, false, true><<<a.B, NT, vsm, stream>>>(a, g, dE); count_launch();
        } else {
            if ((rcv = set_smem(heads_bwd_kernel<32, 1, false, true>, vsm)) != 0) return rcv;
            heads_bwd_kernel<32, 1, false, true><<<a.B, NT, vsm, stream>>>(a, g, dE); count_launch();
        }
        if ((rcv = int(cudaGetLastError())) != 0) return rcv;
        return heads_tc_sweep(a, g, dE, stream);
    }
    const int NF = a.M + 2 * a.K + a.H;
    const int NB = (NF + NT - 1) / NT;
    if (NB > 3 || (NB == 3 && HP != 32)) return FVAE_ERR_LIMIT;
    const bool wsm = HP <= 32 && bwd_smem_bytes(HP, a.H, a.K, a.M, true) <= 110 * 1024;   // keep two CTAs per SM
    const size_t smem = bwd_smem_bytes(HP, a.H, a.K, a.M, wsm);
    int rc;
    // slices of a date's stocks over blockIdx.y: enough CTAs for two per SM, never more slices than 64-stock chunks
    int nsm = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    // one wave: as many slices per date as resident CTAs allow (occupancy of the chosen instantiation, queried below)
    int split = 1;
    const int max_chunks = (int((int64_t(a.S) + a.B - 1) / a.B) + CH - 1) / CH;      // 64-stock chunks of an average date
    auto pick_split = [&](int occ) {
        int sp = (nsm * (occ < 1 ? 1 : occ)) / a.B;
        if (sp > max_chunks) sp = max_chunks;
        if (sp > 16) sp = 16;
        return sp < 1 ? 1 : sp;
    };
#define FVAE_LAUNCH_BWD(HPV, NBV, WSMV)                                                      \
    do {                                                                                     \
        if ((rc = set_smem(heads_bwd_kernel<HPV, NBV, WSMV>, smem)) != 0) return rc;         \
        int occ = 1;                                                                         \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, heads_bwd_kernel<HPV, NBV, WSMV>, NT, smem); \
        split = pick_split(occ);                                                             \
        heads_bwd_kernel<HPV, NBV, WSMV><<<dim3(a.B, split), NT, smem, stream>>>(a, g, dE); count_launch(); \
    } while (0)
    if (HP == 20) {
        if (wsm) { if (NB == 1) FVAE_LAUNCH_BWD(20, 1, true); else FVAE_LAUNCH_BWD(20, 2, true); }
        else { if (NB == 1) FVAE_LAUNCH_BWD(20, 1, false); else FVAE_LAUNCH_BWD(20, 2, false); }
    } else if (HP == 32) {
        if (wsm) { if (NB == 1) FVAE_LAUNCH_BWD(32, 1, true); else if (NB == 2) FVAE_LAUNCH_BWD(32, 2, true); else FVAE_LAUNCH_BWD(32, 3, true); }
        else { if (NB == 1) FVAE_LAUNCH_BWD(32, 1, false); else if (NB == 2) FVAE_LAUNCH_BWD(32, 2, false); else FVAE_LAUNCH_BWD(32, 3, false); }
    } else if (HP == 48) {
        if (NB == 1) FVAE_LAUNCH_BWD(48, 1, false); else FVAE_LAUNCH_BWD(48, 2, false);
    } else {
        if (NB == 1) FVAE_LAUNCH_BWD(64, 1, false); else FVAE_LAUNCH_BWD(64, 2, false);
    }
#undef FVAE_LAUNCH_BWD
    return int(cudaGetLastError());
}

}  // namespace fvae
