// Resident panel: the (date, instrument) row table stays on the device; these kernels replace
// the per-sample pandas/numpy window gather of the reference's TSDataSampler (dataset.py:139-181).
//   fvae_window_index   : _get_indices for a whole batch of samples (look-back rows, ffill / bfill on the ROW NUMBERS,
//                         NaN -> sentinel row) and the label slice returns[:, -1] (train_model.py:18-22)
//   fvae_gather_windows : __getitem__'s data_arr[indices] for callers that want the window tensor itself
// The FeatureExtractor kernels read the rows through the index directly (fvae_panel.row_index), so a training step
// never materialises the T-fold duplicated windows.
#include <cuda_bf16.h>

#include "fvae_common.cuh"

namespace fvae {
namespace {

constexpr int kMaxT = 256;

// one thread per sample: the T row numbers of its look-back window
__global__ void window_index_kernel(const int32_t* __restrict__ idx_mat, int D, int I, const int32_t* __restrict__ sample_date,
                                    const int32_t* __restrict__ sample_inst, int64_t S, int T, int fill_mode, int nan_row,
                                    int32_t* __restrict__ row_index, const float* __restrict__ label, float* __restrict__ y) {
    const int64_t s = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int d = sample_date[s], j = sample_inst[s];
    int32_t* out = row_index + s * T;
    // forward fill (np_ffill, dataset.py:24-39): a missing entry takes the last present one; leading ones stay missing
    int last = -1, first = -1;
    for (int t = 0; t < T; ++t) {
        const int dd = d - T + 1 + t;
        int v = (dd >= 0 && dd < D && j >= 0 && j < I) ? idx_mat[int64_t(dd) * I + j] : -1;
        if (v >= 0 && first < 0) first = v;
        if (fill_mode != FVAE_FILL_NONE) { if (v < 0) v = last; else last = v; }
        out[t] = v;
    }
    // backward fill of the leading gap (ffill of the reversed vector, dataset.py:148), then NaN -> sentinel row (:172)
    int lastv = nan_row;
    for (int t = 0; t < T; ++t) {
        int v = out[t];
        if (v < 0) v = (fill_mode == FVAE_FILL_FFILL_BFILL && first >= 0) ? first : nan_row;
        out[t] = v;
        lastv = v;
    }
    if (label && y) y[s] = label[lastv];
}

template <typename XT, typename OT>
__global__ void gather_windows_kernel(const XT* __restrict__ x, int64_t row_pitch, const int32_t* __restrict__ row_index,
                                      int64_t rows, int C, OT* __restrict__ out) {
    // one warp per (sequence, time) row
    const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const XT* src = x + int64_t(row_index[r]) * row_pitch;
    OT* dst = out + r * C;
    for (int c = lane; c < C; c += 32) dst[c] = OT(float(src[c]));
}

}  // namespace
}  // namespace fvae

using namespace fvae;

extern "C" {

int fvae_window_index(const int32_t* idx_mat, int32_t D, int32_t I, const int32_t* sample_date, const int32_t* sample_inst,
                      int64_t S, int32_t T, int32_t fill_mode, int32_t nan_row, int32_t* row_index, const float* label,
                      float* y, void* stream) {
    if (!idx_mat || !sample_date || !sample_inst || !row_index) return FVAE_ERR_NULL;
    if (D <= 0 || I <= 0 || S <= 0 || T <= 0 || T > kMaxT || nan_row < 0) return FVAE_ERR_SHAPE;
    if (fill_mode < FVAE_FILL_NONE || fill_mode > FVAE_FILL_FFILL_BFILL) return FVAE_ERR_UNSUPPORTED;
    if ((label == nullptr) != (y == nullptr)) return FVAE_ERR_NULL;
    const int threads = 128;
    const int64_t blocks = (S + threads - 1) / threads;
    window_index_kernel<<<unsigned(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        idx_mat, D, I, sample_date, sample_inst, S, T, fill_mode, nan_row, row_index, label, y); count_launch();
    return int(cudaGetLastError());
}

int fvae_gather_windows(const fvae_panel* panel, int64_t S, int32_t T, int32_t C, void* out, int32_t out_dtype, void* stream) {
    if (!panel || !panel->data || !panel->row_index || !out) return FVAE_ERR_NULL;
    if (S <= 0 || T <= 0 || C <= 0 || panel->row_pitch < C || panel->num_rows <= 0) return FVAE_ERR_SHAPE;
    if ((panel->dtype != FVAE_F32 && panel->dtype != FVAE_BF16) || (out_dtype != FVAE_F32 && out_dtype != FVAE_BF16)) return FVAE_ERR_DTYPE;
    const int64_t rows = S * T;
    const int threads = 256;
    const int64_t blocks = (rows * 32 + threads - 1) / threads;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define FVAE_GATHER(XT, OT)                                                                                              \
    gather_windows_kernel<XT, OT><<<unsigned(blocks), threads, 0, st>>>(static_cast<const XT*>(panel->data), panel->row_pitch, \
                                                                     panel->row_index, rows, C, static_cast<OT*>(out))
    if (panel->dtype == FVAE_F32 && out_dtype == FVAE_F32) FVAE_GATHER(float, float);
    else if (panel->dtype == FVAE_F32) FVAE_GATHER(float, __nv_bfloat16);
    else if (out_dtype == FVAE_F32) FVAE_GATHER(__nv_bfloat16, float);
    else FVAE_GATHER(__nv_bfloat16, __nv_bfloat16);
#undef FVAE_GATHER
    count_launch();
    return int(cudaGetLastError());
}

}  // extern "C"
