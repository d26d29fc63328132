// Backward stock sweep of the cross-sectional heads on tcgen05 (FVAE_PREC_BF16_TC only).
//
// The CUDA-core backward (heads.cu) spends its time in three products over the stocks of a date:
//     F  = E . Wcat^T          (E: stocks x H hidden state;  Wcat: stacked rows [Wp; G; Wb; Wa; dp])
//     dE = Z . Wcat            (Z: the per-(stock,row) backward coefficient, a pointwise function of F)
//     dWcat = Z^T . E
// (reference: autograd of module.py:52-67 FactorEncoder, :134-153 AttentionLayer, :107-123 FactorDecoder).
// Here each 128-stock tile of a date runs them as three UMMA groups around one row-local transform:
//     E tile (bf16 hi | lo split, ones column at H folds the biases)
//       -> F in TMEM (hi*hi + lo*hi + hi*lo: fp32-class scores, they feed exp())
//       -> 512 threads read F, write Z (bf16) as a chunk-major tile
//       -> dE (Z K-major x image B2)  and  dW (Z, E both MN-major; accumulators stay in TMEM for the whole CTA).
// The per-date vector phase (KL, mapping layer, predictor head, dy_p, dp_k) stays in heads.cu (VEC mode) and hands
// dy_p / dp_k / pooled_k.dp_k over through the workspace.
//
// Column layout of Wcat / Z (TcCols): [0,M) encoder | Kp attention scores | Kp beta | Hp8 alpha hidden | Kp attention
// weights (rows = dp_k of the date), every group padded to 8 so one thread owns whole 16-byte chunks.
#include <float.h>

#include "heads.cuh"
#include "tc_sm100.cuh"

namespace fvae {

namespace {

using namespace tc;

constexpr int TNT = 512;                       // 4 column parts x 128 rows
constexpr float kL2E = 1.4426950408889634f;
constexpr uint32_t kColDW = 256;               // TMEM columns of the weight-gradient accumulators (2 blocks x 32)
constexpr int kHK = 32;                        // K extent of the E tile: H columns + ones column, padded

__device__ __forceinline__ float sanitize(float v) { return fabsf(v) <= FLT_MAX ? v : 0.f; }

struct RowRef { const float* w; float bias; int kind; int idx; };   // kind: 0 enc, 1 att, 2 beta, 3 alpha, -1 none
__device__ __forceinline__ RowRef row_ref(const HeadsArgs& a, const TcCols& tcg, int c) {
    RowRef r; r.w = nullptr; r.bias = 0.f; r.kind = -1; r.idx = 0;
    const int H = a.H, K = a.K, M = a.M;
    if (c < M) { r.kind = 0; r.idx = c; r.w = a.w.Wp + size_t(c) * H; r.bias = a.w.bp[c]; }
    else if (c >= tcg.c_att && c < tcg.c_att + K) { r.kind = 1; r.idx = c - tcg.c_att; r.w = a.sv.G + size_t(r.idx) * H; r.bias = a.sv.cvec[r.idx]; }
    else if (c >= tcg.c_beta && c < tcg.c_beta + K) { r.kind = 2; r.idx = c - tcg.c_beta; r.w = a.w.Wb + size_t(r.idx) * H; r.bias = a.w.bb[r.idx]; }
    else if (c >= tcg.c_alpha && c < tcg.c_alpha + H) { r.kind = 3; r.idx = c - tcg.c_alpha; r.w = a.w.Wa + size_t(r.idx) * H; r.bias = a.w.ba[r.idx]; }
    return r;
}

// image B1 [8 chunks][NZ rows][8]: chunks 0-3 = bf16 hi of row c (k = h, bias at k = H), chunks 4-7 = lo residual
// image B2 [NZ/8 chunks][32 rows h][8]: B2[h][c] = Wcat[c][h]   (K-major B operand of dE = Z . Wcat)
__global__ void heads_tc_prep_kernel(HeadsArgs a, TcCols tcg) {
    const int H = a.H, NZ = tcg.NZ;
    __nv_bfloat16* b1 = static_cast<__nv_bfloat16*>(a.sv.t_b1);
    __nv_bfloat16* b2 = static_cast<__nv_bfloat16*>(a.sv.t_b2);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < NZ * kHK; idx += gridDim.x * blockDim.x) {
        const int c = idx / kHK, k = idx % kHK;
        const RowRef r = row_ref(a, tcg, c);
        float w = 0.f;
        if (r.kind >= 0) w = (k < H) ? r.w[k] : (k == H ? r.bias : 0.f);
        w = sanitize(w);
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        b1[(size_t(k / 8) * NZ + c) * 8 + k % 8] = hi;
        b1[(size_t(4 + k / 8) * NZ + c) * 8 + k % 8] = lo;
        b2[(size_t(c / 8) * kHK + k) * 8 + c % 8] = (k < H) ? hi : __float2bfloat16_rn(0.f);
    }
}

struct SweepSmem {
    uint32_t zt, et, b1, b2, enc4, att4, bet2, alw, accal, bar, slot, total;
};
__host__ __device__ inline SweepSmem sweep_layout(int H, int K, int M, const TcCols& tcg) {
    SweepSmem s; uint32_t p = 0;
    auto take = [&](uint32_t n) { uint32_t r = p; p += (n + 127u) & ~127u; return r; };
    s.zt = take(32 * kTileChunk);
    s.et = take(8 * kTileChunk);
    s.b1 = take(8u * tcg.NZ * 16u);
    s.b2 = take(uint32_t(tcg.NZ) * 64u);
    s.enc4 = take(uint32_t(M) * 16u);
    s.att4 = take(uint32_t(tcg.Kp) * 16u);
    s.bet2 = take(uint32_t(tcg.Kp) * 8u);
    s.alw = take((2u * 32u + 1u) * 4u);
    s.accal = take((2u * 32u + 2u) * 4u);
    s.bar = take(16);
    s.slot = take(16);
    s.total = p;
    (void)H; (void)K;
    return s;
}

__device__ __forceinline__ void split_store8(uint8_t* hi_dst, uint8_t* lo_dst, const float (&v)[8]) {
    uint32_t ph[4], pl[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * t]), h1 = __float2bfloat16_rn(v[2 * t + 1]);
        const float r0 = v[2 * t] - __bfloat162float(h0), r1 = v[2 * t + 1] - __bfloat162float(h1);
        ph[t] = uint32_t(__bfloat16_as_ushort(h0)) | (uint32_t(__bfloat16_as_ushort(h1)) << 16);
        pl[t] = pack_bf16(r0, r1);
    }
    *reinterpret_cast<uint4*>(hi_dst) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
    *reinterpret_cast<uint4*>(lo_dst) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}
__device__ __forceinline__ void store8(uint8_t* dst, const float (&z)[8]) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16(z[0], z[1]), pack_bf16(z[2], z[3]), pack_bf16(z[4], z[5]), pack_bf16(z[6], z[7]));
}

__global__ void __launch_bounds__(TNT, 1) heads_tc_sweep_kernel(HeadsArgs a, HeadsG g, float* __restrict__ dE, TcCols tcg) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int H = a.H, K = a.K, M = a.M, NZ = tcg.NZ;
    const int tid = threadIdx.x, row = tid & 127, part = tid >> 7, warp = tid >> 5, lane = tid & 31;
    const SweepSmem L = sweep_layout(H, K, M, tcg);
    uint8_t* Zt = smem + L.zt;
    uint8_t* Et = smem + L.et;
    uint8_t* B1 = smem + L.b1;
    uint8_t* B2 = smem + L.b2;
    float4* enc4 = reinterpret_cast<float4*>(smem + L.enc4);   // {-max*log2e, dyp/sum, y_p, -}
    float4* att4 = reinterpret_cast<float4*>(smem + L.att4);   // {max, 1/sum, pooled.dp, guard}
    float2* bet2 = reinterpret_cast<float2*>(smem + L.bet2);   // {mu_z, sigma_z^2}
    float* alw = reinterpret_cast<float*>(smem + L.alw);       // wam[32] | was[32] | bas
    float* accal = reinterpret_cast<float*>(smem + L.accal);   // d wam[32] | d was[32] | d bam | d bas
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + L.slot);
    __shared__ int s_red[TNT / 32];

    if (warp == 0) tmem_alloc<512>(slot);
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    {   // static images, zeroed Z tile, alpha-layer vectors
        const uint4* s1 = static_cast<const uint4*>(a.sv.t_b1);
        for (int i = tid; i < 8 * NZ; i += TNT) reinterpret_cast<uint4*>(B1)[i] = s1[i];
        const uint4* s2 = static_cast<const uint4*>(a.sv.t_b2);
        for (int i = tid; i < NZ * 4; i += TNT) reinterpret_cast<uint4*>(B2)[i] = s2[i];
        for (int i = tid; i < 32 * 128; i += TNT) reinterpret_cast<uint4*>(Zt)[i] = make_uint4(0, 0, 0, 0);
        for (int j = tid; j < 32; j += TNT) {
            alw[j] = (j < H) ? a.w.wam[j] : 0.f;
            alw[32 + j] = (j < H) ? a.w.was[j] : 0.f;
        }
        if (tid == 0) alw[64] = a.w.bas[0];
        for (int j = tid; j < 66; j += TNT) accal[j] = 0.f;
    }
    // my contiguous range of 128-stock tiles
    int total = 0;
    {
        int cnt = 0;
        for (int d = tid; d < a.B; d += TNT) { const int n = a.date_ptr[d + 1] - a.date_ptr[d]; cnt += n > 0 ? (n + 127) / 128 : 0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) s_red[warp] = cnt;
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    for (int w = 0; w < TNT / 32; ++w) total += s_red[w];
    const uint32_t tmem = *slot;
    const uint32_t lane_base = tmem + (uint32_t((warp & 3) * 32) << 16);
    const int per = (total + int(gridDim.x) - 1) / int(gridDim.x);
    const int lo = int(blockIdx.x) * per, hi = min(total, lo + per);

    const uint32_t zt_addr = smem_u32(Zt), et_addr = smem_u32(Et), b1_addr = smem_u32(B1), b2_addr = smem_u32(B2);
    const float inv_tau = 1.f / sqrtf(float(H) + 1e-6f);
    const int nblk = (NZ + 127) / 128;
    uint32_t ph = 0;
    bool first = true;

    int gt = 0;
    for (int d = 0; d < a.B && gt < hi; ++d) {
        const int p0 = a.date_ptr[d], n = a.date_ptr[d + 1] - p0;
        const int nt = n > 0 ? (n + 127) / 128 : 0;
        if (nt == 0 || gt + nt <= lo) { gt += nt; continue; }
        const int t0 = max(0, lo - gt), t1 = min(nt, hi - gt);
        gt += nt;
        const float coefN = 2.f / (float(n) * float(a.B));
        // ---- per-date vectors and the dp rows of both images (the previous tile ended with a barrier)
        for (int j = tid; j < M; j += TNT)
            enc4[j] = make_float4(-a.sv.enc_m[size_t(d) * M + j] * kL2E, a.sv.t_dyp[size_t(d) * M + j] / a.sv.enc_l[size_t(d) * M + j],
                                  a.sv.yp[size_t(d) * M + j], 0.f);
        for (int k = tid; k < tcg.Kp; k += TNT) {
            float4 v = make_float4(0.f, 0.f, 0.f, 1.f);
            float2 b = make_float2(0.f, 0.f);
            if (k < K) {
                const size_t o = size_t(d) * K + k;
                v = make_float4(a.sv.att_m[o], 1.f / a.sv.att_l[o], a.sv.t_pdp[o], a.sv.bad[o] ? 1.f : 0.f);
                const float sg = a.out.sigma_post[o];
                b = make_float2(a.out.mu_post[o], sg * sg);
            }
            att4[k] = v; bet2[k] = b;
        }
        for (int idx = tid; idx < tcg.Kp * kHK; idx += TNT) {
            const int k = idx / kHK, h = idx % kHK;
            float v = 0.f;
            if (k < K && h < H && !a.sv.bad[size_t(d) * K + k]) v = sanitize(a.sv.t_dps[(size_t(d) * K + k) * H + h]);
            const __nv_bfloat16 vh = __float2bfloat16_rn(v);
            const __nv_bfloat16 vl = __float2bfloat16_rn(v - __bfloat162float(vh));
            const int c = tcg.c_atta + k;
            reinterpret_cast<__nv_bfloat16*>(B1)[(size_t(h / 8) * NZ + c) * 8 + h % 8] = vh;
            reinterpret_cast<__nv_bfloat16*>(B1)[(size_t(4 + h / 8) * NZ + c) * 8 + h % 8] = vl;
            reinterpret_cast<__nv_bfloat16*>(B2)[(size_t(c / 8) * kHK + h) * 8 + c % 8] = vh;
        }
        for (int t = t0; t < t1; ++t) {
            const int i = t * 128 + row;
            const bool valid = i < n;
            const int u = p0 + i;
            // ---- E tile: part p stages columns [8p, 8p+8) of my row, hi | lo
            {
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int h = part * 8 + q;
                    v[q] = (valid && h < H) ? a.e[size_t(u) * H + h] : ((valid && h == H) ? 1.f : 0.f);
                }
                split_store8(Et + tile_off(128, row, part), Et + tile_off(128, row, 4 + part), v);
            }
            fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after_sync();
                const uint32_t idesc = make_idesc_bf16(kTileRows, uint32_t(NZ), false, false);
                const uint32_t bch = uint32_t(NZ) * 16u;
                auto mm = [&](uint32_t ac, uint32_t bc, uint32_t acc) {
                    mma_bf16_ss(tmem, make_smem_desc(et_addr + ac * kTileChunk, kTileChunk, 128),
                                make_smem_desc(b1_addr + bc * bch, bch, 128), idesc, acc);
                };
                mm(0, 0, 0); mm(2, 2, 1);      // hi . hi
                mm(4, 0, 1); mm(6, 2, 1);      // lo . hi
                mm(0, 4, 1); mm(2, 6, 1);      // hi . lo
                mma_commit(bar);
            }
            mbar_wait(bar, ph); ph ^= 1;
            tc_fence_after_sync();
            // ---- transform F -> Z
            if (part < 2) {                                   // encoder rows: w_ij dy_p_j (y_i - y_p_j)
                const int nch = M / 8, half = (nch + 1) / 2;
                const int c0 = part == 0 ? 0 : half, c1 = part == 0 ? half : nch;
                const float yi = valid ? a.y[u] : 0.f;
                for (int ch = c0; ch < c1; ++ch) {
                    float f[8], z[8];
                    tmem_ld8(lane_base + uint32_t(ch * 8), f);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 e4 = enc4[ch * 8 + q];
                        z[q] = valid ? exp2f(fmaf(f[q], kL2E, e4.x)) * e4.y * (yi - e4.z) : 0.f;
                    }
                    store8(Zt + tile_off(128, row, ch), z);
                }
            } else if (part == 2) {                           // attention: d score and the weights a_ik
                for (int kc = 0; kc < tcg.Kp / 8; ++kc) {
                    float fs[8], fa[8], zs[8], za[8];
                    tmem_ld8(lane_base + uint32_t(tcg.c_att + kc * 8), fs);
                    tmem_ld8(lane_base + uint32_t(tcg.c_atta + kc * 8), fa);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int k = kc * 8 + q;
                        const float4 a4 = att4[k];
                        zs[q] = 0.f; za[q] = 0.f;
                        if (valid && a4.w == 0.f) {
                            const float kf = keep_factor(a, u, k);
                            const float x = fs[q] * inv_tau * kf;
                            const float aik = expf(relu_nan(x) - a4.x) * a4.y;
                            za[q] = aik;
                            zs[q] = (x > 0.f) ? kf * inv_tau * aik * (fa[q] - a4.z) : 0.f;
                        }
                    }
                    store8(Zt + tile_off(128, row, tcg.c_att / 8 + kc), zs);
                    store8(Zt + tile_off(128, row, tcg.c_atta / 8 + kc), za);
                }
            } else {                                          // beta rows, alpha hidden rows and the alpha scalars
                float v1 = 0.f, v2 = 0.f;                     // d loss / d mu_y, d loss / d sigma_y^2
                if (valid) {
                    v1 = coefN * (a.out.yhat[u] - a.y[u]);
                    v2 = v1 * eps_of(a, u) / (2.f * a.out.sigma_y[u]);
                }
                for (int kc = 0; kc < tcg.Kp / 8; ++kc) {
                    float f[8], z[8];
                    tmem_ld8(lane_base + uint32_t(tcg.c_beta + kc * 8), f);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float2 b = bet2[kc * 8 + q];
                        z[q] = b.x * v1 + 2.f * f[q] * b.y * v2;
                    }
                    store8(Zt + tile_off(128, row, tcg.c_beta / 8 + kc), z);
                }
                float hp[32];
#pragma unroll
                for (int jc = 0; jc < 4; ++jc) {
                    float f[8];
                    if (jc < tcg.Hp8 / 8) tmem_ld8(lane_base + uint32_t(tcg.c_alpha + jc * 8), f);
#pragma unroll
                    for (int q = 0; q < 8; ++q) hp[jc * 8 + q] = (jc < tcg.Hp8 / 8) ? f[q] : 0.f;
                }
                float asp = alw[64];
#pragma unroll
                for (int j = 0; j < 32; ++j) asp = fmaf(alw[32 + j], lrelu(hp[j]), asp);
                const float asig = softplus(asp);
                const float damu = v1;
                const float dasp = valid ? 2.f * asig * v2 * softplus_grad(asp) : 0.f;
#pragma unroll
                for (int jc = 0; jc < 4; ++jc) {
                    if (jc < tcg.Hp8 / 8) {
                        float z[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int j = jc * 8 + q;
                            z[q] = (damu * alw[j] + dasp * alw[32 + j]) * (hp[j] > 0.f ? 1.f : kLeakySlope);
                        }
                        store8(Zt + tile_off(128, row, tcg.c_alpha / 8 + jc), z);
                    }
                }
                // mu / sigma layer gradients: column sums over the 32 rows of this warp, then shared accumulators
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j < H) {
                        const float ha = valid ? lrelu(hp[j]) : 0.f;
                        float s1 = damu * ha, s2 = dasp * ha;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                        }
                        if (lane == 0) { atomicAdd(accal + j, s1); atomicAdd(accal + 32 + j, s2); }
                    }
                }
                {
                    float s1 = damu, s2 = dasp;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                    }
                    if (lane == 0) { atomicAdd(accal + 64, s1); atomicAdd(accal + 65, s2); }
                }
            }
            fence_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after_sync();
                issue_row_gemm_acc(tmem, 0, zt_addr, b2_addr, kHK, kHK, NZ / 16, false);              // dE -> columns [0,32)
                for (int b = 0; b < nblk; ++b)
                    issue_wgrad_acc(tmem, kColDW + 32u * b, zt_addr, 16u * b, et_addr, kHK, !first);  // dWcat block b
                mma_commit(bar);
            }
            first = false;
            mbar_wait(bar, ph); ph ^= 1;
            tc_fence_after_sync();
            {   // dE out: part p owns hidden columns [8p, 8p+8)
                float v[8];
                tmem_ld8(lane_base + uint32_t(part * 8), v);
                if (valid) {
                    float* dst = dE + size_t(u) * H;
#pragma unroll
                    for (int q = 0; q < 8; ++q) { const int h = part * 8 + q; if (h < H) dst[h] = v[q]; }
                }
            }
            tc_fence_before_sync();
            __syncthreads();
        }
    }
    // ---- flush the weight-gradient accumulators: TMEM lane = stacked row c, column = h (bias at h = H)
    if (!first) {
        if (part < nblk) {
            float w[32];
            {
                float t8[8];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    tmem_ld8(lane_base + kColDW + 32u * uint32_t(part) + 8u * q4, t8);
#pragma unroll
                    for (int q = 0; q < 8; ++q) w[q4 * 8 + q] = t8[q];
                }
            }
            const int c = part * 128 + row;
            const RowRef r = row_ref(a, tcg, c);
            float* gw = nullptr; float* gb = nullptr;
            if (r.kind == 0) { gw = g.Wp + size_t(r.idx) * H; gb = g.bp + r.idx; }
            else if (r.kind == 1) { gw = a.sv.dG + size_t(r.idx) * H; gb = a.sv.dc + r.idx; }
            else if (r.kind == 2) { gw = g.Wb + size_t(r.idx) * H; gb = g.bb + r.idx; }
            else if (r.kind == 3) { gw = g.Wa + size_t(r.idx) * H; gb = g.ba + r.idx; }
            if (gw) {
#pragma unroll
                for (int h = 0; h < 32; ++h) {
                    if (h < H) atomicAdd(gw + h, w[h]);
                    else if (h == H) atomicAdd(gb, w[h]);
                }
            }
        }
        for (int j = tid; j < H; j += TNT) { atomicAdd(g.wam + j, accal[j]); atomicAdd(g.was + j, accal[32 + j]); }
        if (tid == 0) { atomicAdd(g.bam, accal[64]); atomicAdd(g.bas, accal[65]); }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

}  // namespace

int64_t heads_tc_image_bytes(int H, int K, int M, int which) {
    const TcCols c = tc_cols(H, K, M);
    return which == 1 ? int64_t(8) * c.NZ * 16 : int64_t(c.NZ) * 64;
}

int heads_tc_prep(const HeadsArgs& a, cudaStream_t stream) {
    const TcCols c = tc_cols(a.H, a.K, a.M);
    heads_tc_prep_kernel<<<(c.NZ * kHK + 255) / 256, 256, 0, stream>>>(a, c); count_launch();
    return int(cudaGetLastError());
}

int heads_tc_sweep(const HeadsArgs& a, const HeadsG& g, float* dE, cudaStream_t stream) {
    const TcCols c = tc_cols(a.H, a.K, a.M);
    const SweepSmem L = sweep_layout(a.H, a.K, a.M, c);
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    cudaError_t e = cudaFuncSetAttribute(heads_tc_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(L.total));
    if (e != cudaSuccess) return int(e);
    int rc = heads_tc_prep(a, stream);
    if (rc != 0) return rc;
    heads_tc_sweep_kernel<<<sms, TNT, L.total, stream>>>(a, g, dE, c); count_launch();
    return int(cudaGetLastError());
}

}  // namespace fvae
