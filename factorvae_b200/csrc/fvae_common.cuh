// Shared host/device helpers of the FactorVAE hot path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>

#include "../../include/fvae_b200.h"

namespace fvae {

constexpr float kLeakySlope = 0.01f;        // nn.LeakyReLU default   (reference module.py:19,73,164)
constexpr float kLnEps = 1e-5f;             // nn.LayerNorm default   (module.py:17)
constexpr float kSoftplusThreshold = 20.f;  // nn.Softplus default    (module.py:42,76,167)
constexpr float kKeepScale = 1.0f / 0.9f;   // nn.Dropout(0.1)        (module.py:132)
constexpr float kSigmaFloor = 1e-6f;        // module.py:117, :265
constexpr int kMaxC = 192;
constexpr int kMaxH = 64;

// diagnostics only: number of kernel launches issued by this library (bench.py reports it)
extern unsigned long long g_launch_count;
inline void count_launch(int n = 1) { __atomic_fetch_add(&g_launch_count, (unsigned long long)n, __ATOMIC_RELAXED); }

// ---- flat parameter layout --------------------------------------------------------------
struct Layout {
    int64_t off[FVAE_P_NUM_SECTIONS + 1];
};

inline int64_t align4(int64_t v) { return (v + 3) & ~int64_t(3); }

inline Layout make_layout(int C, int H, int K, int M) {
    int64_t n[FVAE_P_NUM_SECTIONS];
    n[FVAE_P_LN_W] = C;            n[FVAE_P_LN_B] = C;
    n[FVAE_P_W1] = int64_t(C) * C; n[FVAE_P_B1] = C;
    n[FVAE_P_WIH] = int64_t(3) * H * C; n[FVAE_P_WHH] = int64_t(3) * H * H;
    n[FVAE_P_BIH] = 3 * H;         n[FVAE_P_BHH] = 3 * H;
    n[FVAE_P_ENC_W] = int64_t(M) * H; n[FVAE_P_ENC_B] = M;
    n[FVAE_P_ENC_MU_W] = int64_t(K) * M; n[FVAE_P_ENC_MU_B] = K;
    n[FVAE_P_ENC_SG_W] = int64_t(K) * M; n[FVAE_P_ENC_SG_B] = K;
    n[FVAE_P_AL_W] = int64_t(H) * H; n[FVAE_P_AL_B] = H;
    n[FVAE_P_AL_MU_W] = H; n[FVAE_P_AL_MU_B] = 1; n[FVAE_P_AL_SG_W] = H; n[FVAE_P_AL_SG_B] = 1;
    n[FVAE_P_BETA_W] = int64_t(K) * H; n[FVAE_P_BETA_B] = K;
    n[FVAE_P_ATT_Q] = int64_t(K) * H;
    n[FVAE_P_ATT_KW] = int64_t(K) * H * H; n[FVAE_P_ATT_KB] = int64_t(K) * H;
    n[FVAE_P_ATT_VW] = int64_t(K) * H * H; n[FVAE_P_ATT_VB] = int64_t(K) * H;
    n[FVAE_P_PR_W] = int64_t(H) * H; n[FVAE_P_PR_B] = H;
    n[FVAE_P_PR_MU_W] = H; n[FVAE_P_PR_MU_B] = 1; n[FVAE_P_PR_SG_W] = H; n[FVAE_P_PR_SG_B] = 1;
    Layout L;
    int64_t o = 0;
    for (int i = 0; i < FVAE_P_NUM_SECTIONS; ++i) { L.off[i] = o; o = align4(o + n[i]); }
    L.off[FVAE_P_NUM_SECTIONS] = o;
    return L;
}

// ---- device helpers ---------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : kLeakySlope * v; }
__device__ __forceinline__ float lrelu_grad_from_out(float out) { return out > 0.f ? 1.f : kLeakySlope; }
// torch.nn.Softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplus(float v) { return v > kSoftplusThreshold ? v : log1pf(expf(v)); }
__device__ __forceinline__ float softplus_grad(float v) { return v > kSoftplusThreshold ? 1.f : 1.f / (1.f + expf(-v)); }
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum; `scratch` holds >= 32 floats.  All threads get the result.
// vector reduction into global memory (16-byte aligned address): a quarter of the atomic traffic of four scalar adds
__device__ __forceinline__ void red_add_v4(float* addr, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

__device__ __forceinline__ float block_sum(float v, float* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    float r = (lane < nw) ? scratch[lane] : 0.f;
    r = warp_sum(r);
    return r;
}

// ---- Philox4x32-10 counter RNG (shard-invariant noise: keyed by global unit / head) ------
__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__device__ __forceinline__ uint4 philox4(uint64_t seed, uint64_t ctr_lo, uint64_t ctr_hi) {
    uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
    uint32_t c0 = uint32_t(ctr_lo), c1 = uint32_t(ctr_lo >> 32), c2 = uint32_t(ctr_hi), c3 = uint32_t(ctr_hi >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float u32_to_unit(uint32_t v) { return (float(v >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)
// eps for global unit g at training step `step`: Box-Muller on one Philox draw.
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t step, int64_t g) {
    uint4 r = philox4(seed, uint64_t(g), (step << 1));
    float u1 = u32_to_unit(r.x), u2 = u32_to_unit(r.y);
    return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}
// keep decision (prob 0.9) for (unit g, head k): a counter hash (two murmur-style finalizer rounds over
// (seed, step, unit, head)) -- dropout masks do not need Philox strength and this is evaluated once per
// (stock, head) in forward and again in backward.  Same shard-invariance contract as the eps stream.
__device__ __forceinline__ bool philox_keep(uint64_t seed, uint64_t step, int64_t g, int k) {
    uint32_t h = uint32_t(g) * 0x9E3779B1u ^ (uint32_t(uint64_t(g) >> 32) * 0x7FEB352Du) ^ (uint32_t(k) * 0x85EBCA77u)
               ^ uint32_t(seed) ^ (uint32_t(seed >> 32) * 0xC2B2AE3Du) ^ (uint32_t(step) * 0x27D4EB2Fu);
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    h += uint32_t(k) * 0x9E3779B9u + uint32_t(g);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h >= 0x1999999Au;          // 2^32 / 10
}
#endif  // __CUDACC__

}  // namespace fvae
