// Diagnostics: the step's in-kernel noise, evaluated stand-alone with the SAME device functions the
// heads kernels call (fvae_common.cuh: philox_normal / philox_keep), so a parity test can replay a
// FVAE_FLAG_PHILOX step through the CPU oracle with the kernel's own eps (reference module.py:104)
// and dropout keep decisions (module.py:132,144), and can measure the keep rate of the counter hash.
#include "fvae_common.cuh"

namespace fvae {
namespace {

__global__ void noise_debug_kernel(uint64_t seed, uint64_t step, int64_t unit_base, int64_t S, int K, float* eps,
                                   uint8_t* keep) {
    const int64_t nth = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < S; i += nth) {
        const int64_t g = unit_base + i;
        if (eps) eps[i] = philox_normal(seed, step, g);
        if (keep)
            for (int k = 0; k < K; ++k) keep[i * K + k] = philox_keep(seed, step, g, k) ? 1 : 0;
    }
}

}  // namespace
}  // namespace fvae

extern "C" int fvae_debug_noise(uint64_t seed, uint64_t step, int64_t unit_base, int64_t S, int32_t K, float* eps,
                                uint8_t* keep_mask, void* stream) {
    if (S <= 0 || K <= 0) return FVAE_ERR_SHAPE;
    if (!eps && !keep_mask) return FVAE_ERR_NULL;
    const int grid = int((S + 255) / 256 < 4096 ? (S + 255) / 256 : 4096);
    fvae::noise_debug_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(seed, step, unit_base, S, K, eps, keep_mask);
    fvae::count_launch();
    return int(cudaGetLastError());
}
