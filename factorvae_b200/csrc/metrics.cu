// Per-date Spearman rank correlation of predictions and labels -- the reference's RankIC
// (utils.py:113-129): for every date, pandas `.rank()` (average ranks for ties) of both columns, scipy.stats.spearmanr of
// the ranks, then mean / std / IR over the dates on the host.  Here: one CTA per date, bitonic sort in shared memory,
// average ranks, Pearson correlation of the ranks.  Any NaN in a date, fewer than 2 stocks or a constant column -> NaN,
// like pandas + scipy.
#include <float.h>
#include <math.h>

#include "fvae_common.cuh"

namespace fvae {
namespace {

constexpr int kRankMaxN = 4096;       // stocks per date held in shared memory
constexpr int kRankThreads = 256;

// sort (key, idx) ascending by the TOTAL order (key, idx); n2 = power of two >= n, padded with (+inf, idx >= n).  The
// tie-break on idx matters: a bitonic network is not stable, and with a plain key compare the padding slots could land in
// front of a real +inf (or NaN -> +inf) element, i.e. at a sorted position p < n, and be ranked as if they were stocks.
__device__ void bitonic_sort(float* key, int* idx, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const float a = key[i], b = key[l];
                    const bool gt = a > b || (a == b && idx[i] > idx[l]);
                    if (gt == up) { key[i] = b; key[l] = a; const int t = idx[i]; idx[i] = idx[l]; idx[l] = t; }
                }
            }
            __syncthreads();
        }
    }
}

// average ranks (1-based) of v[0..n) into rank[0..n); returns through *has_nan whether a NaN was seen
__device__ void average_ranks(const float* __restrict__ v, int n, int n2, float* key, int* idx, float* rank, int* has_nan) {
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        float x = INFINITY;
        if (i < n) { x = v[i]; if (x != x) { *has_nan = 1; x = INFINITY; } }
        key[i] = x; idx[i] = i;
    }
    __syncthreads();
    bitonic_sort(key, idx, n2);
    for (int p = threadIdx.x; p < n; p += blockDim.x) {     // positions [0, n) hold exactly the real entries (total order)
        const float x = key[p];
        int lo = p, hi = p;
        while (lo > 0 && key[lo - 1] == x) --lo;
        while (hi + 1 < n && key[hi + 1] == x) ++hi;
        rank[idx[p]] = 0.5f * float(lo + hi) + 1.f;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kRankThreads) rank_ic_kernel(const float* __restrict__ pred, const float* __restrict__ label,
                                                               const int32_t* __restrict__ date_ptr, float* __restrict__ ric) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ float red[32];
    __shared__ int has_nan;
    const int d = blockIdx.x;
    const int p0 = date_ptr[d], n = date_ptr[d + 1] - p0;
    if (n < 2 || n > kRankMaxN) { if (threadIdx.x == 0) ric[d] = nanf(""); return; }
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    float* key = reinterpret_cast<float*>(smem);
    int* idx = reinterpret_cast<int*>(key + n2);
    float* ra = reinterpret_cast<float*>(idx + n2);
    float* rb = ra + n;
    if (threadIdx.x == 0) has_nan = 0;
    __syncthreads();
    average_ranks(pred + p0, n, n2, key, idx, ra, &has_nan);
    average_ranks(label + p0, n, n2, key, idx, rb, &has_nan);
    const float mu = 0.5f * float(n + 1);
    float sab = 0.f, saa = 0.f, sbb = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float a = ra[i] - mu, b = rb[i] - mu;
        sab = fmaf(a, b, sab); saa = fmaf(a, a, saa); sbb = fmaf(b, b, sbb);
    }
    sab = block_sum(sab, red);
    saa = block_sum(saa, red);
    sbb = block_sum(sbb, red);
    if (threadIdx.x == 0) ric[d] = (has_nan || saa == 0.f || sbb == 0.f) ? nanf("") : sab / sqrtf(saa * sbb);
}

}  // namespace
}  // namespace fvae

using namespace fvae;

extern "C" int fvae_rank_ic(const float* pred, const float* label, const int32_t* date_ptr, int32_t B, int32_t max_per_date,
                            float* ric, void* stream) {
    if (!pred || !label || !date_ptr || !ric) return FVAE_ERR_NULL;
    if (B <= 0 || max_per_date <= 0) return FVAE_ERR_SHAPE;
    if (max_per_date > kRankMaxN) return FVAE_ERR_LIMIT;
    int n2 = 1;
    while (n2 < max_per_date) n2 <<= 1;
    const size_t smem = size_t(n2) * 8 + size_t(max_per_date) * 8;
    cudaError_t e = cudaFuncSetAttribute(rank_ic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return int(e);
    rank_ic_kernel<<<B, kRankThreads, smem, static_cast<cudaStream_t>(stream)>>>(pred, label, date_ptr, ric); count_launch();
    return int(cudaGetLastError());
}
