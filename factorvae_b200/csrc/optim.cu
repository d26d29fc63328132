// Fused optimizer step: torch.optim.Adam over the FLAT parameter / gradient buffers
// (reference main.py:60 `torch.optim.Adam(factorVAE.parameters(), lr=args.lr)`, stepped per batch in
// train_model.py:30; defaults betas (0.9, 0.999), eps 1e-8, weight_decay 0, amsgrad off).  One launch instead of the
// 28 + 5K per-tensor updates; same arithmetic as torch's single-tensor Adam:
//     g' = g * grad_scale (+ weight_decay * p);  m = lerp(m, g', 1 - beta1);  v = beta2 v + (1 - beta2) g'^2
//     p -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
#include <math.h>

#include "fvae_common.cuh"

namespace fvae {
namespace {

__global__ void adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                 int64_t n, float lr_over_bc1, float beta1, float beta2, float eps, float weight_decay,
                                 float inv_sqrt_bc2, float grad_scale) {
    const int64_t i4 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    float pv[4], gv[4], mv[4], vv[4];
    const bool full = i4 + 4 <= n;
    if (full) {
        const float4 a = *reinterpret_cast<const float4*>(p + i4), b = *reinterpret_cast<const float4*>(g + i4);
        const float4 c = *reinterpret_cast<const float4*>(m + i4), d = *reinterpret_cast<const float4*>(v + i4);
        pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w; gv[0] = b.x; gv[1] = b.y; gv[2] = b.z; gv[3] = b.w;
        mv[0] = c.x; mv[1] = c.y; mv[2] = c.z; mv[3] = c.w; vv[0] = d.x; vv[1] = d.y; vv[2] = d.z; vv[3] = d.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool ok = i4 + e < n;
            pv[e] = ok ? p[i4 + e] : 0.f; gv[e] = ok ? g[i4 + e] : 0.f; mv[e] = ok ? m[i4 + e] : 0.f; vv[e] = ok ? v[i4 + e] : 0.f;
        }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float gr = gv[e] * grad_scale;
        if (weight_decay != 0.f) gr = fmaf(weight_decay, pv[e], gr);
        mv[e] = fmaf(1.f - beta1, gr - mv[e], mv[e]);                    // exp_avg.lerp_(grad, 1 - beta1)
        vv[e] = fmaf(1.f - beta2, gr * gr, beta2 * vv[e]);               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = sqrtf(vv[e]) * inv_sqrt_bc2 + eps;
        pv[e] = pv[e] - lr_over_bc1 * (mv[e] / denom);
    }
    if (full) {
        *reinterpret_cast<float4*>(p + i4) = make_float4(pv[0], pv[1], pv[2], pv[3]);
        *reinterpret_cast<float4*>(m + i4) = make_float4(mv[0], mv[1], mv[2], mv[3]);
        *reinterpret_cast<float4*>(v + i4) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (i4 + e < n) { p[i4 + e] = pv[e]; m[i4 + e] = mv[e]; v[i4 + e] = vv[e]; }
    }
}

}  // namespace
}  // namespace fvae

using namespace fvae;

extern "C" int fvae_adam_step(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                              void* stream) {
    if (!params || !grad || !exp_avg || !exp_avg_sq) return FVAE_ERR_NULL;
    if (n <= 0 || step <= 0) return FVAE_ERR_SHAPE;
    if ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15u) return FVAE_ERR_SHAPE;            // float4 accesses
    const double bc1 = 1.0 - pow(double(beta1), double(step));
    const double bc2 = 1.0 - pow(double(beta2), double(step));
    const float lr_over_bc1 = float(double(lr) / bc1);
    const float inv_sqrt_bc2 = float(1.0 / sqrt(bc2));
    const int threads = 256;
    const int64_t blocks = ((n + 3) / 4 + threads - 1) / threads;
    adam_step_kernel<<<unsigned(blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        params, grad, exp_avg, exp_avg_sq, n, lr_over_bc1, beta1, beta2, eps, weight_decay, inv_sqrt_bc2, grad_scale);
    count_launch();
    return int(cudaGetLastError());
}
