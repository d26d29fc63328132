// FeatureExtractor, FVAE_PREC_FP32: CUDA-core fp32 kernels (tight-parity mode, sm_100a).
//
// Restates reference module.py:22-31 (LayerNorm :26, Linear+LeakyReLU :27-28, GRU :30, h_T :31)
// and its autograd.  This is the accuracy mode (ELBO within 1e-5 of the reference fp32 path); the
// throughput mode is fe_tc.cu.  Structure: the rows (sequence, time) are processed in chunks that
// fit a scratch area, with three generic kernels
//     rowgemm : OUT[r][n] = epi(sum_k A[r][k] W(k,n) + bias[n])       (x-proj, du, dxn)
//     wgrad   : OUT[a][b] += sum_r A[r][a] B[r][b]                     (dW1, dW_ih, dW_hh)
//     colprod : out[n] += sum_r A[r][n] * B[r][n]                      (bias / LayerNorm grads)
// around a per-sequence GRU recurrence kernel and its BPTT counterpart.
#include "fe.cuh"

namespace fvae {
namespace {

constexpr int kChunkRows = 32768;

struct WsF32 {
    float *gi, *hall, *dgh, *xhat, *xn, *u, *dpre, *dxn;
    int64_t bytes;
};

WsF32 carve_f32(const FeDims& d, void* base) {
    WsF32 w;
    char* p = static_cast<char*>(base);
    auto take = [&](int64_t n) { float* r = reinterpret_cast<float*>(p); p += ((n * 4 + 255) / 256) * 256; return r; };
    const int64_t R = int64_t(d.S) * d.T;
    const int64_t rc = R < kChunkRows ? R : kChunkRows;
    w.gi = take(R * 3 * d.H);
    w.hall = take(R * d.H);
    w.dgh = take(R * 3 * d.H);
    w.xhat = take(rc * d.C);
    w.xn = take(rc * d.C);
    w.u = take(rc * d.C);
    w.dpre = take(rc * d.C);
    w.dxn = take(rc * d.C);
    w.bytes = p - static_cast<char*>(base);
    return w;
}

// ---- LayerNorm over the C features of each (sequence, time) row: one warp per row ------------
template <typename XT>
__global__ void ln_rows_kernel(const XT* __restrict__ x, int64_t seq_pitch, int64_t row_pitch, const int32_t* __restrict__ row_index,
                               int T, int C, int64_t row0, int nrows, const float* __restrict__ gamma,
                               const float* __restrict__ beta, float* __restrict__ xn, float* __restrict__ xhat) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    const int64_t row = row0 + warp;
    const XT* src = row_index ? x + int64_t(row_index[row]) * row_pitch : x + (row / T) * seq_pitch + (row % T) * row_pitch;
    float v[kMaxC / 32];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxC / 32; ++j) {
        const int c = lane + 32 * j;
        v[j] = (c < C) ? float(src[c]) : 0.f;
        sum += v[j];
    }
    const float mean = warp_sum(sum) / float(C);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxC / 32; ++j) {
        const int c = lane + 32 * j;
        const float dlt = (c < C) ? v[j] - mean : 0.f;
        sq = fmaf(dlt, dlt, sq);
    }
    const float rstd = rsqrtf(warp_sum(sq) / float(C) + kLnEps);
#pragma unroll
    for (int j = 0; j < kMaxC / 32; ++j) {
        const int c = lane + 32 * j;
        if (c < C) {
            const float xh = (v[j] - mean) * rstd;
            xn[size_t(warp) * C + c] = fmaf(xh, gamma[c], beta[c]);
            if (xhat) xhat[size_t(warp) * C + c] = xh;
        }
    }
}

// ---- rowgemm: 32-row tile, all N (<=192) columns, K streamed in chunks of 32 --------------------
// W is addressed as W[n*ldw + k] (w_nk = 1, an nn.Linear weight used forward) or W[k*ldw + n].
enum { EPI_NONE = 0, EPI_LRELU = 1, EPI_MUL_LRELU_GRAD = 2 };

__global__ void __launch_bounds__(256) rowgemm_kernel(const float* __restrict__ A, int lda, int R, int Kd,
                                                       const float* __restrict__ W, int ldw, int w_nk, int N,
                                                       const float* __restrict__ bias, int epi,
                                                       const float* __restrict__ aux, float* __restrict__ out, int ldo) {
    __shared__ float As[32][33];
    __shared__ float Ws[32][193];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.x * 32;
    float acc[4][6];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < Kd; k0 += 32) {
        __syncthreads();
        for (int idx = tid; idx < 32 * 32; idx += 256) {
            const int r = idx >> 5, kk = idx & 31;
            As[r][kk] = (r0 + r < R && k0 + kk < Kd) ? A[size_t(r0 + r) * lda + k0 + kk] : 0.f;
        }
        if (w_nk) {
            for (int idx = tid; idx < N * 32; idx += 256) {
                const int n = idx >> 5, kk = idx & 31;
                Ws[kk][n] = (k0 + kk < Kd) ? W[size_t(n) * ldw + k0 + kk] : 0.f;
            }
        } else {
            for (int idx = tid; idx < 32 * N; idx += 256) {
                const int kk = idx / N, n = idx % N;
                Ws[kk][n] = (k0 + kk < Kd) ? W[size_t(k0 + kk) * ldw + n] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < 32; ++kk) {
            float a[4], b[6];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[4 * warp + i][kk];
#pragma unroll
            for (int j = 0; j < 6; ++j) b[j] = (lane + 32 * j < N) ? Ws[kk][lane + 32 * j] : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = r0 + 4 * warp + i;
        if (r >= R) continue;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int n = lane + 32 * j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (epi == EPI_LRELU) v = lrelu(v);
            else if (epi == EPI_MUL_LRELU_GRAD) v *= lrelu_grad_from_out(aux[size_t(r) * ldo + n]);
            out[size_t(r) * ldo + n] = v;
        }
    }
}

// ---- wgrad: OUT[a][b] += sum_r A[r][a] B[r][b]; rows split over blockIdx.y ------------------------
// period > 0: rows with (r % period) == skip_phase are skipped (the t=0 rows of dW_hh).
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ A, int lda, int na,
                                                     const float* __restrict__ Bm, int ldb, int nb, int64_t R,
                                                     int rows_per_cta, int period, int skip_phase,
                                                     float* __restrict__ out, int ldo) {
    __shared__ float As[32][33];
    __shared__ float Bs[32][193];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int a0 = blockIdx.x * 32;
    const int64_t rbeg = int64_t(blockIdx.y) * rows_per_cta;
    const int64_t rend = (rbeg + rows_per_cta < R) ? rbeg + rows_per_cta : R;
    float acc[4][6];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) acc[i][j] = 0.f;
    for (int64_t rr0 = rbeg; rr0 < rend; rr0 += 32) {
        __syncthreads();
        for (int idx = tid; idx < 32 * 32; idx += 256) {
            const int r = idx >> 5, aa = idx & 31;
            const int64_t row = rr0 + r;
            const bool ok = row < rend && a0 + aa < na && !(period > 0 && (row % period) == skip_phase);
            As[r][aa] = ok ? A[row * lda + a0 + aa] : 0.f;
        }
        for (int idx = tid; idx < 32 * nb; idx += 256) {
            const int r = idx / nb, b = idx % nb;
            const int64_t row = rr0 + r;
            Bs[r][b] = (row < rend) ? Bm[row * ldb + b] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
            float a[4], b[6];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[r][4 * warp + i];
#pragma unroll
            for (int j = 0; j < 6; ++j) b[j] = (lane + 32 * j < nb) ? Bs[r][lane + 32 * j] : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 6; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int aa = a0 + 4 * warp + i;
        if (aa >= na) continue;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int b = lane + 32 * j;
            if (b < nb) atomicAdd(out + size_t(aa) * ldo + b, acc[i][j]);
        }
    }
}

// ---- colprod: out_prod[n] += sum_r A[r][n]*B[r][n] ; out_sum[n] += sum_r A[r][n] -------------------
__global__ void colprod_kernel(const float* __restrict__ A, const float* __restrict__ Bm, int ld, int N, int64_t R,
                               int rows_per_cta, float* __restrict__ out_prod, float* __restrict__ out_sum) {
    const int n = threadIdx.x;
    if (n >= N) return;
    const int64_t rbeg = int64_t(blockIdx.x) * rows_per_cta;
    const int64_t rend = (rbeg + rows_per_cta < R) ? rbeg + rows_per_cta : R;
    float s1 = 0.f, s2 = 0.f;
    for (int64_t r = rbeg; r < rend; ++r) {
        const float av = A[r * ld + n];
        s2 += av;
        if (Bm) s1 = fmaf(av, Bm[r * ld + n], s1);
    }
    if (out_prod) atomicAdd(out_prod + n, s1);
    if (out_sum) atomicAdd(out_sum + n, s2);
}

// ---- GRU recurrence (module.py:30): 4 sequences per CTA, thread = (sequence, hidden unit) ----------
//   r = sig(gi_r + W_hr h + b_hr); z = sig(gi_z + W_hz h + b_hz); n = tanh(gi_n + r*(W_hn h + b_hn))
//   h' = (1 - z) n + z h            (gi already holds W_i* u + b_i*)
constexpr int GSEQ = 4;

__global__ void __launch_bounds__(64 * GSEQ) gru_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ Whh,
                                                            const float* __restrict__ bhh, int S, int T, int H,
                                                            float* __restrict__ hall, float* __restrict__ e) {
    extern __shared__ float sm[];
    float* Wt = sm;                         // [H][3H]   Wt[k][g] = Whh[g][k]
    float* hs = Wt + H * 3 * H;             // [2][GSEQ][64]
    const int j = threadIdx.x, sl = threadIdx.y;
    const int tid = sl * 64 + j;
    for (int idx = tid; idx < 3 * H * H; idx += 64 * GSEQ) {
        const int g = idx / H, k = idx % H;
        Wt[k * 3 * H + g] = Whh[idx];
    }
    hs[sl * 64 + j] = 0.f;
    hs[GSEQ * 64 + sl * 64 + j] = 0.f;
    __syncthreads();
    const int s = blockIdx.x * GSEQ + sl;
    const bool act = s < S && j < H;
    const float br = act ? bhh[j] : 0.f, bz = act ? bhh[H + j] : 0.f, bn = act ? bhh[2 * H + j] : 0.f;
    int cur = 0;
    float h = 0.f;
    for (int t = 0; t < T; ++t) {
        const float* hp = hs + cur * GSEQ * 64 + sl * 64;
        float ar = br, az = bz, an = bn;
        if (act) {
            for (int k = 0; k < H; ++k) {
                const float hk = hp[k];
                const float* wr = Wt + k * 3 * H;
                ar = fmaf(wr[j], hk, ar);
                az = fmaf(wr[H + j], hk, az);
                an = fmaf(wr[2 * H + j], hk, an);
            }
            const float* g3 = gi + (size_t(s) * T + t) * 3 * H;
            const float r = sigmoidf_(g3[j] + ar);
            const float z = sigmoidf_(g3[H + j] + az);
            const float n = tanhf(g3[2 * H + j] + r * an);
            h = (1.f - z) * n + z * h;
            hall[(size_t(s) * T + t) * H + j] = h;
            hs[(cur ^ 1) * GSEQ * 64 + sl * 64 + j] = h;
        }
        cur ^= 1;
        __syncthreads();
    }
    if (act) e[size_t(s) * H + j] = h;
}

// ---- GRU BPTT: dgi written over gi, dgh to its own buffer ---------------------------------------
//   dn = dh (1-z); dz = dh (h_prev - n); da_n = dn (1-n^2); da_r = da_n hn r(1-r); da_z = dz z(1-z)
//   dgi = [da_r, da_z, da_n]; dgh = [da_r, da_z, da_n r]; dh_prev = dh z + dgh . W_hh
__global__ void __launch_bounds__(64 * GSEQ) gru_bwd_kernel(float* __restrict__ gi, const float* __restrict__ hall,
                                                            const float* __restrict__ dE,
                                                            const float* __restrict__ Whh, const float* __restrict__ bhh,
                                                            int S, int T, int H, float* __restrict__ dgh) {
    extern __shared__ float sm[];
    float* Wt = sm;                         // [H][3H]
    float* Wr = Wt + H * 3 * H;             // [3H][H]  (row-major copy)
    float* hps = Wr + 3 * H * H;            // [GSEQ][64]   h_{t-1}
    float* dgs = hps + GSEQ * 64;           // [GSEQ][3*64] dgh of this step
    const int j = threadIdx.x, sl = threadIdx.y;
    const int tid = sl * 64 + j;
    for (int idx = tid; idx < 3 * H * H; idx += 64 * GSEQ) {
        const int g = idx / H, k = idx % H;
        const float v = Whh[idx];
        Wt[k * 3 * H + g] = v;
        Wr[idx] = v;
    }
    const int s = blockIdx.x * GSEQ + sl;
    const bool act = s < S && j < H;
    const float br = act ? bhh[j] : 0.f, bz = act ? bhh[H + j] : 0.f, bn = act ? bhh[2 * H + j] : 0.f;
    float dh = act ? dE[size_t(s) * H + j] : 0.f;
    for (int t = T - 1; t >= 0; --t) {
        __syncthreads();
        const float hprev = (act && t > 0) ? hall[(size_t(s) * T + t - 1) * H + j] : 0.f;
        hps[sl * 64 + j] = hprev;
        __syncthreads();
        float dar = 0.f, daz = 0.f, dan = 0.f, z = 0.f, r = 0.f;
        if (act) {
            float ar = br, az = bz, an = bn;
            const float* hp = hps + sl * 64;
            for (int k = 0; k < H; ++k) {
                const float hk = hp[k];
                const float* wr = Wt + k * 3 * H;
                ar = fmaf(wr[j], hk, ar);
                az = fmaf(wr[H + j], hk, az);
                an = fmaf(wr[2 * H + j], hk, an);
            }
            float* g3 = gi + (size_t(s) * T + t) * 3 * H;
            r = sigmoidf_(g3[j] + ar);
            z = sigmoidf_(g3[H + j] + az);
            const float n = tanhf(g3[2 * H + j] + r * an);
            const float dn = dh * (1.f - z);
            const float dz = dh * (hprev - n);
            dan = dn * (1.f - n * n);
            dar = dan * an * r * (1.f - r);
            daz = dz * z * (1.f - z);
            g3[j] = dar; g3[H + j] = daz; g3[2 * H + j] = dan;
            float* q3 = dgh + (size_t(s) * T + t) * 3 * H;
            q3[j] = dar; q3[H + j] = daz; q3[2 * H + j] = dan * r;
            dgs[sl * 192 + j] = dar; dgs[sl * 192 + H + j] = daz; dgs[sl * 192 + 2 * H + j] = dan * r;
        }
        __syncthreads();
        if (act) {
            float acc = dh * z;
            const float* dg = dgs + sl * 192;
            for (int g = 0; g < 3 * H; ++g) acc = fmaf(dg[g], Wr[g * H + j], acc);
            dh = acc;
        }
    }
}

inline int grid_rows(int64_t rows, int per) { return int((rows + per - 1) / per); }

template <typename XT>
void launch_ln(const fvae_panel& x, const FeDims& d, int64_t row0, int nrows, const FeW& w, float* xn, float* xhat,
               cudaStream_t st) {
    const int warps_per_cta = 8;
    ln_rows_kernel<XT><<<grid_rows(nrows, warps_per_cta), 32 * warps_per_cta, 0, st>>>(
        static_cast<const XT*>(x.data), x.seq_pitch, x.row_pitch, x.row_index, d.T, d.C, row0, nrows, w.ln_w, w.ln_b, xn, xhat); count_launch();
}

void run_ln(const fvae_panel& x, const FeDims& d, int64_t row0, int nrows, const FeW& w, float* xn, float* xhat,
            cudaStream_t st) {
    if (x.dtype == FVAE_BF16) launch_ln<__nv_bfloat16>(x, d, row0, nrows, w, xn, xhat, st);
    else launch_ln<float>(x, d, row0, nrows, w, xn, xhat, st);
}

}  // namespace

int64_t fe_f32_workspace_bytes(const FeDims& d) { return carve_f32(d, nullptr).bytes; }

// shared with fe_tc.cu while its recurrence / backward still run the fp32 kernels
struct FeF32Views { float *gi, *hall; };
FeF32Views fe_f32_views(const FeDims& d, void* ws) {
    WsF32 W = carve_f32(d, ws);
    return FeF32Views{W.gi, W.hall};
}
int fe_f32_gru_forward(const FeDims& d, const FeW& w, void* ws, float* e, cudaStream_t st) {
    WsF32 W = carve_f32(d, ws);
    const size_t smem = (size_t(3) * d.H * d.H + 2 * GSEQ * 64) * sizeof(float);
    cudaError_t ce = cudaFuncSetAttribute(gru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (ce != cudaSuccess) return int(ce);
    gru_fwd_kernel<<<grid_rows(d.S, GSEQ), dim3(64, GSEQ), smem, st>>>(W.gi, w.Whh, w.bhh, d.S, d.T, d.H, W.hall, e); count_launch();
    return int(cudaGetLastError());
}

int fe_f32_forward(const FeDims& d, const fvae_panel& x, const FeW& w, float* e, void* ws, cudaStream_t st) {
    WsF32 W = carve_f32(d, ws);
    const int64_t R = int64_t(d.S) * d.T;
    const int H3 = 3 * d.H;
    for (int64_t r0 = 0; r0 < R; r0 += kChunkRows) {
        const int nr = int((R - r0 < kChunkRows) ? R - r0 : kChunkRows);
        run_ln(x, d, r0, nr, w, W.xn, nullptr, st);
        rowgemm_kernel<<<grid_rows(nr, 32), 256, 0, st>>>(W.xn, d.C, nr, d.C, w.W1, d.C, 1, d.C, w.b1, EPI_LRELU, nullptr,
                                                          W.u, d.C); count_launch();
        rowgemm_kernel<<<grid_rows(nr, 32), 256, 0, st>>>(W.u, d.C, nr, d.C, w.Wih, d.C, 1, H3, w.bih, EPI_NONE, nullptr,
                                                          W.gi + r0 * H3, H3); count_launch();
    }
    const size_t smem = (size_t(3) * d.H * d.H + 2 * GSEQ * 64) * sizeof(float);
    cudaError_t ce = cudaFuncSetAttribute(gru_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (ce != cudaSuccess) return int(ce);
    gru_fwd_kernel<<<grid_rows(d.S, GSEQ), dim3(64, GSEQ), smem, st>>>(W.gi, w.Whh, w.bhh, d.S, d.T, d.H, W.hall, e); count_launch();
    return int(cudaGetLastError());
}

int fe_f32_backward(const FeDims& d, const fvae_panel& x, const FeW& w, const FeG& g, const float* dE, void* ws,
                    cudaStream_t st) {
    WsF32 W = carve_f32(d, ws);
    const int64_t R = int64_t(d.S) * d.T;
    const int H3 = 3 * d.H, H = d.H, C = d.C;
    // BPTT: gi -> dgi (in place), dgh
    {
        const size_t smem = (size_t(6) * H * H + GSEQ * 64 + GSEQ * 192) * sizeof(float);
        cudaError_t ce = cudaFuncSetAttribute(gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (ce != cudaSuccess) return int(ce);
        gru_bwd_kernel<<<grid_rows(d.S, GSEQ), dim3(64, GSEQ), smem, st>>>(W.gi, W.hall, dE, w.Whh, w.bhh, d.S, d.T, H,
                                                                         W.dgh); count_launch();
    }
    const int rows_per_cta = 2048;
    // dW_hh[g][k] = sum_{s,t>0} dgh[s,t,g] h[s,t-1,k] : pair dgh row j+1 with hall row j, skipping j % T == T-1
    if (R > 1 && d.T > 1)
        wgrad_kernel<<<dim3(grid_rows(H3, 32), grid_rows(R - 1, rows_per_cta)), 256, 0, st>>>(
            W.dgh + H3, H3, H3, W.hall, H, H, R - 1, rows_per_cta, d.T, d.T - 1, g.Whh, H); count_launch();
    colprod_kernel<<<grid_rows(R, rows_per_cta), 192, 0, st>>>(W.dgh, nullptr, H3, H3, R, rows_per_cta, nullptr, g.bhh); count_launch();
    colprod_kernel<<<grid_rows(R, rows_per_cta), 192, 0, st>>>(W.gi, nullptr, H3, H3, R, rows_per_cta, nullptr, g.bih); count_launch();
    // front: recompute xn / u per row chunk, then the three weight gradients and LayerNorm's
    for (int64_t r0 = 0; r0 < R; r0 += kChunkRows) {
        const int nr = int((R - r0 < kChunkRows) ? R - r0 : kChunkRows);
        const float* dgi = W.gi + r0 * H3;
        run_ln(x, d, r0, nr, w, W.xn, W.xhat, st);
        rowgemm_kernel<<<grid_rows(nr, 32), 256, 0, st>>>(W.xn, C, nr, C, w.W1, C, 1, C, w.b1, EPI_LRELU, nullptr, W.u, C); count_launch();
        // du = dgi . W_ih ; dpre = du * lrelu'(pre)
        rowgemm_kernel<<<grid_rows(nr, 32), 256, 0, st>>>(dgi, H3, nr, H3, w.Wih, C, 0, C, nullptr, EPI_MUL_LRELU_GRAD, W.u,
                                                          W.dpre, C); count_launch();
        // dxn = dpre . W1
        rowgemm_kernel<<<grid_rows(nr, 32), 256, 0, st>>>(W.dpre, C, nr, C, w.W1, C, 0, C, nullptr, EPI_NONE, nullptr, W.dxn,
                                                          C); count_launch();
        const int ysplit = grid_rows(nr, rows_per_cta);
        wgrad_kernel<<<dim3(grid_rows(H3, 32), ysplit), 256, 0, st>>>(dgi, H3, H3, W.u, C, C, nr, rows_per_cta, 0, 0, g.Wih, C); count_launch();
        wgrad_kernel<<<dim3(grid_rows(C, 32), ysplit), 256, 0, st>>>(W.dpre, C, C, W.xn, C, C, nr, rows_per_cta, 0, 0, g.W1, C); count_launch();
        colprod_kernel<<<ysplit, 192, 0, st>>>(W.dpre, nullptr, C, C, nr, rows_per_cta, nullptr, g.b1); count_launch();
        colprod_kernel<<<ysplit, 192, 0, st>>>(W.dxn, W.xhat, C, C, nr, rows_per_cta, g.ln_w, g.ln_b); count_launch();
    }
    return int(cudaGetLastError());
}

}  // namespace fvae
