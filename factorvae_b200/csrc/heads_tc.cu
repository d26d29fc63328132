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

constexpr int TNT = 512;                       // transform threads: 4 column parts x 128 rows
constexpr int TNT_ALL = TNT + 32;              // + one warp that only issues the UMMAs
constexpr float kL2E = 1.4426950408889634f;
constexpr int kHK = 32;                        // K extent of the E tile: H columns + ones column, padded

__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
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

// image B1s [8 chunks][NS rows][8]: chunks 0-3 = bf16 hi of static row c (k = h, bias at k = H), chunks 4-7 = lo residual
// image B2s [NS/8 chunks][32 rows h][8]: B2s[h][c] = Wcat[c][h]   (K-major B operand of dE = Z . Wcat)
// block 0 also writes the prefix of 128-stock tiles per date (tile_ptr[B+1]).
__global__ void heads_tc_prep_kernel(HeadsArgs a, TcCols tcg) {
    const int H = a.H, NS = tcg.NS;
    __nv_bfloat16* b1 = static_cast<__nv_bfloat16*>(a.sv.t_b1);
    __nv_bfloat16* b2 = static_cast<__nv_bfloat16*>(a.sv.t_b2);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < NS * kHK; idx += gridDim.x * blockDim.x) {
        const int c = idx / kHK, k = idx % kHK;
        const RowRef r = row_ref(a, tcg, c);
        float w = 0.f;
        if (r.kind >= 0) w = (k < H) ? r.w[k] : (k == H ? r.bias : 0.f);
        w = sanitize(w);
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
        b1[(size_t(k / 8) * NS + c) * 8 + k % 8] = hi;
        b1[(size_t(4 + k / 8) * NS + c) * 8 + k % 8] = lo;
        b2[(size_t(c / 8) * kHK + k) * 8 + c % 8] = (k < H) ? hi : __float2bfloat16_rn(0.f);
    }
    if (blockIdx.x == 0) {
        __shared__ int part[256];
        const int B = a.B, tid = threadIdx.x, q = (B + 255) / 256;
        int cnt = 0;
        for (int d = tid * q; d < min(B, (tid + 1) * q); ++d) { const int n = a.date_ptr[d + 1] - a.date_ptr[d]; cnt += n > 0 ? (n + 127) / 128 : 0; }
        part[tid] = cnt;
        __syncthreads();
        if (tid == 0) { int run = 0; for (int i = 0; i < 256; ++i) { const int c = part[i]; part[i] = run; run += c; } a.sv.t_tile_ptr[B] = run; }
        __syncthreads();
        int run = part[tid];
        for (int d = tid * q; d < min(B, (tid + 1) * q); ++d) {
            a.sv.t_tile_ptr[d] = run;
            const int n = a.date_ptr[d + 1] - a.date_ptr[d];
            run += n > 0 ? (n + 127) / 128 : 0;
        }
    }
}

struct SweepSmem {
    uint32_t zt, et, b1s, b2s, b1d, b2d, enc4, att4, bet2, alw, accal, bar, slot, start, total;
};
constexpr uint32_t kB1dBytes = 8u * 32u * 16u;     // per-date image: dp rows, hi | lo      [8 chunks][32 rows k][8]
constexpr uint32_t kB2dBytes = 4u * 32u * 16u;     // per-date image: B2d[h][k] = dp_k[h]    [4 chunks][32 rows h][8]
__host__ __device__ inline SweepSmem sweep_layout(int M, const TcCols& tcg) {
    SweepSmem s; uint32_t p = 0;
    auto take = [&](uint32_t n) { uint32_t r = p; p += (n + 127u) & ~127u; return r; };
    s.zt = take(32 * kTileChunk);
    s.et = take(3 * 8 * kTileChunk);
    s.b1s = take(8u * tcg.NS * 16u);
    s.b2s = take(uint32_t(tcg.NS) * 64u);
    s.b1d = take(2 * kB1dBytes);
    s.b2d = take(2 * kB2dBytes);
    s.enc4 = take(2u * uint32_t(M) * 16u);
    s.att4 = take(2u * 32u * 16u);
    s.bet2 = take(2u * 32u * 8u);
    s.alw = take((2u * 32u + 1u) * 4u);
    s.accal = take(64u * 4u);
    s.bar = take(32);
    s.slot = take(16);
    s.start = take(16);
    s.total = p;
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

// Sum each of 32 per-lane values over the warp with 31 shuffles; lane l returns the total of v[l].
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float keep = up ? v[i + o] : v[i];
            const float send = up ? v[i] : v[i + o];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

// TMEM columns: F [0, NS+32) | dE double buffer | weight-gradient accumulators (2 blocks x 32)
constexpr uint32_t kColDE = 256, kColDW = 320;

struct TileIt { int d, t, p0, n; };   // date, tile inside the date, first unit of the date, stocks of the date

__global__ void __launch_bounds__(TNT_ALL, 1) heads_tc_sweep_kernel(HeadsArgs a, HeadsG g, float* __restrict__ dE, TcCols tcg) {
    const uint64_t nstep = noise_step(a);
    extern __shared__ __align__(1024) uint8_t smem[];
    const int H = a.H, K = a.K, M = a.M, NS = tcg.NS;
    const int tid = threadIdx.x, row = tid & 127, part = tid >> 7, warp = tid >> 5, lane = tid & 31;
    const SweepSmem L = sweep_layout(M, tcg);
    uint8_t* Zt = smem + L.zt;
    uint8_t* Et = smem + L.et;                                 // 3 buffers of 8 chunks (hi | lo)
    uint8_t* B1s = smem + L.b1s;
    uint8_t* B2s = smem + L.b2s;
    uint8_t* B1d = smem + L.b1d;                               // 2 buffers (date parity)
    uint8_t* B2d = smem + L.b2d;
    float4* enc4 = reinterpret_cast<float4*>(smem + L.enc4);   // [2][M]  {-max*log2e, dyp/sum, y_p, -}
    float4* att4 = reinterpret_cast<float4*>(smem + L.att4);   // [2][32] {max, 1/sum, pooled.dp, guard}
    float2* bet2 = reinterpret_cast<float2*>(smem + L.bet2);   // [2][32] {mu_z, sigma_z^2}
    float* alw = reinterpret_cast<float*>(smem + L.alw);       // wam[32] | was[32] | bas
    float* accal = reinterpret_cast<float*>(smem + L.accal);   // [0,32): d wam (31: d bam) | [32,64): d was (63: d bas)
    uint64_t* barA = reinterpret_cast<uint64_t*>(smem + L.bar);
    uint64_t* barB = barA + 1;                                 // MMA groups complete -> transform threads
    uint64_t* zA = barA + 2;                                   // encoder half of Z (and the next E tile) written -> issuer
    uint64_t* zB = barA + 3;
    const bool issuer = warp == TNT / 32;
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + L.slot);
    int* s_start = reinterpret_cast<int*>(smem + L.start);

    if (warp == 0) tmem_alloc<512>(slot);
    if (tid == 0) { mbar_init(barA, 1); mbar_init(barB, 1); mbar_init(zA, TNT); mbar_init(zB, TNT); mbar_fence_init(); }
    // my contiguous range of 128-stock tiles
    const int total = a.sv.t_tile_ptr[a.B];
    const int per = (total + int(gridDim.x) - 1) / int(gridDim.x);
    const int lo = int(blockIdx.x) * per, hi = min(total, lo + per);
    if (lo < hi)
        for (int d = tid; d < a.B; d += TNT_ALL) {
            const int t0 = a.sv.t_tile_ptr[d], t1 = a.sv.t_tile_ptr[d + 1];
            if (t0 <= lo && lo < t1) { s_start[0] = d; s_start[1] = lo - t0; }
        }
    {   // static images, zeroed Z tile, alpha-layer vectors
        const uint4* s1 = static_cast<const uint4*>(a.sv.t_b1);
        for (int i = tid; i < 8 * NS; i += TNT_ALL) reinterpret_cast<uint4*>(B1s)[i] = s1[i];
        const uint4* s2 = static_cast<const uint4*>(a.sv.t_b2);
        for (int i = tid; i < NS * 4; i += TNT_ALL) reinterpret_cast<uint4*>(B2s)[i] = s2[i];
        for (int i = tid; i < 32 * 128; i += TNT_ALL) reinterpret_cast<uint4*>(Zt)[i] = make_uint4(0, 0, 0, 0);
        for (int j = tid; j < 32; j += TNT_ALL) {
            alw[j] = (j < H) ? a.w.wam[j] : 0.f;
            alw[32 + j] = (j < H) ? a.w.was[j] : 0.f;
        }
        if (tid == 0) alw[64] = a.w.bas[0];
        for (int j = tid; j < 64; j += TNT_ALL) accal[j] = 0.f;
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *slot;
    const uint32_t lane_base = tmem + (uint32_t((warp & 3) * 32) << 16);
    const uint32_t zt_addr = smem_u32(Zt), et_addr = smem_u32(Et), b1s_addr = smem_u32(B1s), b2s_addr = smem_u32(B2s);
    const uint32_t b1d_addr = smem_u32(B1d), b2d_addr = smem_u32(B2d);
    const float inv_tau = 1.f / sqrtf(float(H) + 1e-6f);
    const int ntiles = hi - lo;

    auto load_date = [&](TileIt& it) {
        it.p0 = a.date_ptr[it.d];
        it.n = a.date_ptr[it.d + 1] - it.p0;
    };
    auto advance = [&](TileIt& it) {                         // next tile in date order (skipping empty dates)
        if ((it.t + 1) * 128 < it.n) { ++it.t; return; }
        it.t = 0;
        do { ++it.d; if (it.d >= a.B) { it.n = 0; return; } load_date(it); } while (it.n <= 0);
    };
    // per-date vectors and the dp rows of the per-date images, into parity buffer pb
    auto stage_date = [&](int d, int pb) {
        float4* e4 = enc4 + pb * M;
        for (int j = tid; j < M; j += TNT)
            e4[j] = make_float4(-a.sv.enc_m[size_t(d) * M + j] * kL2E, a.sv.t_dyp[size_t(d) * M + j] / a.sv.enc_l[size_t(d) * M + j],
                                a.sv.yp[size_t(d) * M + j], 0.f);
        for (int k = tid; k < 32; k += TNT) {
            float4 v = make_float4(0.f, 0.f, 0.f, 1.f);
            float2 b = make_float2(0.f, 0.f);
            if (k < K) {
                const size_t o = size_t(d) * K + k;
                v = make_float4(a.sv.att_m[o], 1.f / a.sv.att_l[o], a.sv.t_pdp[o], a.sv.bad[o] ? 1.f : 0.f);
                const float sg = a.out.sigma_post[o];
                b = make_float2(a.out.mu_post[o], sg * sg);
            }
            att4[pb * 32 + k] = v; bet2[pb * 32 + k] = b;
        }
        __nv_bfloat16* i1 = reinterpret_cast<__nv_bfloat16*>(B1d + pb * kB1dBytes);
        __nv_bfloat16* i2 = reinterpret_cast<__nv_bfloat16*>(B2d + pb * kB2dBytes);
        for (int idx = tid; idx < 32 * kHK; idx += TNT) {
            const int k = idx / kHK, h = idx % kHK;
            float v = 0.f;
            if (k < K && h < H && !a.sv.bad[size_t(d) * K + k]) v = sanitize(a.sv.t_dps[(size_t(d) * K + k) * H + h]);
            const __nv_bfloat16 vh = __float2bfloat16_rn(v);
            const __nv_bfloat16 vl = __float2bfloat16_rn(v - __bfloat162float(vh));
            i1[(size_t(h / 8) * 32 + k) * 8 + h % 8] = vh;
            i1[(size_t(4 + h / 8) * 32 + k) * 8 + h % 8] = vl;
            i2[(size_t(k / 8) * 32 + h) * 8 + k % 8] = vh;
        }
    };
    auto load_e = [&](const TileIt& it, float (&v)[8]) {     // part p: columns [8p, 8p+8) of my row (ones column at H)
        const int i = it.t * 128 + row;
        const bool valid = i < it.n;
        const float* src = a.e + size_t(it.p0 + i) * H;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int h = part * 8 + q;
            v[q] = (valid && h < H) ? src[h] : ((valid && h == H) ? 1.f : 0.f);
        }
    };
    auto store_e = [&](int buf, const float (&v)[8]) {
        uint8_t* eb = Et + buf * 8 * kTileChunk;
        split_store8(eb + tile_off(128, row, part), eb + tile_off(128, row, 4 + part), v);
    };
    // F columns [dcol, dcol+N) = E(buf) . image rows^T with the hi/lo split: hi.hi + lo.hi + hi.lo
    auto issue_f = [&](int buf, uint32_t img_addr, uint32_t img_rows, uint32_t row0, uint32_t N, uint32_t dcol) {
        const uint32_t idesc = make_idesc_bf16(kTileRows, N, false, false);
        const uint32_t ea = et_addr + buf * 8 * kTileChunk, bch = img_rows * 16u, ba = img_addr + row0 * 16u;
        auto mm = [&](uint32_t ac, uint32_t bc, uint32_t acc) {
            mma_bf16_ss(tmem + dcol, make_smem_desc(ea + ac * kTileChunk, kTileChunk, 128), make_smem_desc(ba + bc * bch, bch, 128), idesc, acc);
        };
        mm(0, 0, 0); mm(2, 2, 1); mm(4, 0, 1); mm(6, 2, 1); mm(0, 4, 1); mm(2, 6, 1);
    };
    auto store_de = [&](const TileIt& it, int buf) {         // part p owns hidden columns [8p, 8p+8)
        float v[8];
        tmem_ld8(lane_base + kColDE + 32u * buf + uint32_t(part * 8), v);
        const int i = it.t * 128 + row;
        if (i < it.n) {
            float* dst = dE + size_t(it.p0 + i) * H;
#pragma unroll
            for (int q = 0; q < 8; ++q) { const int h = part * 8 + q; if (h < H) dst[h] = v[q]; }
        }
    };

    uint32_t phA = 0, phB = 0;
    if (ntiles > 0) {
        TileIt cur; cur.d = s_start[0]; cur.t = s_start[1]; load_date(cur);
        int pb = 0;                                          // date parity buffer of `cur`
        if (!issuer) {
            stage_date(cur.d, pb);
            float v[8]; load_e(cur, v); store_e(0, v);
        }
        fence_async_smem();
        __syncthreads();
        if (issuer) {
            // ---- UMMA issuer warp: waits for the halves of Z, issues dE / dW of this tile and F of the next one
            if (lane == 0) {
                tc_fence_after_sync();
                issue_f(0, b1s_addr, NS, 0, 128, 0);
                mma_commit(barA);
                issue_f(0, b1s_addr, NS, 128, uint32_t(NS - 128), 128);
                issue_f(0, b1d_addr + pb * kB1dBytes, 32, 0, 32, uint32_t(NS));
                mma_commit(barB);
                uint32_t pzA = 0, pzB = 0;
                for (int g_i = 0; g_i < ntiles; ++g_i) {
                    const bool has_next = g_i + 1 < ntiles;
                    TileIt nxt = cur;
                    if (has_next) advance(nxt);
                    const int pbn = (has_next && nxt.d != cur.d) ? (pb ^ 1) : pb;
                    const uint32_t e_cur = et_addr + (g_i % 3) * 8 * kTileChunk;
                    const uint32_t decol = kColDE + 32u * (g_i & 1);
                    mbar_wait(zA, pzA); pzA ^= 1;
                    tc_fence_after_sync();
                    issue_row_gemm_acc(tmem, decol, zt_addr, b2s_addr, kHK, kHK, 8, false);              // dE: encoder part
                    issue_wgrad_acc(tmem, kColDW, zt_addr, 0, e_cur, kHK, g_i > 0);
                    if (has_next) issue_f((g_i + 1) % 3, b1s_addr, NS, 0, 128, 0);
                    mma_commit(barA);
                    mbar_wait(zB, pzB); pzB ^= 1;
                    tc_fence_after_sync();
                    issue_row_gemm_acc(tmem, decol, zt_addr + 16 * kTileChunk, b2s_addr + 16 * (kHK * 16), kHK, kHK, (NS - 128) / 16, true);
                    issue_row_gemm_acc(tmem, decol, zt_addr + uint32_t(NS / 8) * kTileChunk, b2d_addr + pb * kB2dBytes, kHK, kHK, 2, true);
                    issue_wgrad_acc(tmem, kColDW + 32u, zt_addr, 16, e_cur, kHK, g_i > 0);
                    if (has_next) {
                        issue_f((g_i + 1) % 3, b1s_addr, NS, 128, uint32_t(NS - 128), 128);
                        issue_f((g_i + 1) % 3, b1d_addr + pbn * kB1dBytes, 32, 0, 32, uint32_t(NS));
                    }
                    mma_commit(barB);
                    cur = nxt; pb = pbn;
                }
            }
            __syncwarp();
        } else {
        TileIt prev = cur;
        for (int g_i = 0; g_i < ntiles; ++g_i) {
            const bool has_next = g_i + 1 < ntiles;
            TileIt nxt = cur;
            if (has_next) advance(nxt);
            const bool new_date = has_next && nxt.d != cur.d;
            const int pbn = new_date ? (pb ^ 1) : pb;
            float en[8];
            if (has_next) load_e(nxt, en);
            const int i = cur.t * 128 + row;
            const bool valid = i < cur.n;
            const int u = cur.p0 + i;
            const float4* e4v = enc4 + pb * M;
            // ---- phase A: encoder columns  z = w_ij dy_p_j (y_i - y_p_j)
            mbar_wait(barA, phA); phA ^= 1;
            tc_fence_after_sync();
            {
                const float yi = valid ? a.y[u] : 0.f;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int ch = part * 4 + c4;
                    float f[8], z[8];
                    tmem_ld8(lane_base + uint32_t(ch * 8), f);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 e4 = e4v[ch * 8 + q];
                        z[q] = valid ? ex2_fast(fmaf(f[q], kL2E, e4.x)) * e4.y * (yi - e4.z) : 0.f;
                    }
                    store8(Zt + tile_off(128, row, ch), z);
                }
            }
            if (has_next) store_e((g_i + 1) % 3, en);
            fence_async_smem();
            tc_fence_before_sync();
            mbar_arrive(zA);
            // ---- phase B: attention, beta, alpha columns
            mbar_wait(barB, phB); phB ^= 1;
            tc_fence_after_sync();
            if (g_i > 0) store_de(prev, (g_i - 1) & 1);
            if (new_date) {                                   // next date's vectors: visible to every transform thread before its phase A
                stage_date(nxt.d, pbn);
                asm volatile("bar.sync 1, %0;" ::"n"(TNT) : "memory");
            }
            if (part < 3) {                                   // attention: d score and the weights a_ik
                const float4* a4v = att4 + pb * 32;
                for (int kc = part; kc < tcg.Kp / 8; kc += 3) {
                    float fs[8], fa[8], zs[8], za[8];
                    tmem_ld8(lane_base + uint32_t(tcg.c_att + kc * 8), fs);
                    tmem_ld8(lane_base + uint32_t(tcg.c_atta + kc * 8), fa);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int k = kc * 8 + q;
                        const float4 a4 = a4v[k];
                        zs[q] = 0.f; za[q] = 0.f;
                        if (valid && a4.w == 0.f) {
                            const float kf = keep_factor(a, nstep, u, k);
                            const float x = fs[q] * inv_tau * kf;
                            const float aik = ex2_fast((relu_nan(x) - a4.x) * kL2E) * a4.y;     // same form as the forward (ex2.approx: 2 ulp)
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
                    const float coefN = 2.f / (float(cur.n) * float(a.B));
                    v1 = coefN * (a.out.yhat[u] - a.y[u]);
                    v2 = v1 * eps_of(a, nstep, u) / (2.f * a.out.sigma_y[u]);
                }
                const float2* b2v = bet2 + pb * 32;
                for (int kc = 0; kc < tcg.Kp / 8; ++kc) {
                    float f[8], z[8];
                    tmem_ld8(lane_base + uint32_t(tcg.c_beta + kc * 8), f);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float2 b = b2v[kc * 8 + q];
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
                // mu / sigma layer gradients: column sums over the 32 rows of this warp (H <= 31: slot 31 carries the bias grad)
                float r[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) { hp[j] = lrelu(hp[j]); r[j] = (j < 31) ? damu * hp[j] : damu; }
                const float s1 = warp_transpose_sum(r, lane);
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = (j < 31) ? dasp * hp[j] : dasp;
                const float s2 = warp_transpose_sum(r, lane);
                atomicAdd(accal + lane, s1);
                atomicAdd(accal + 32 + lane, s2);
            }
            fence_async_smem();
            tc_fence_before_sync();
            mbar_arrive(zB);
            prev = cur; cur = nxt; pb = pbn;
        }
        mbar_wait(barB, phB); phB ^= 1;
        tc_fence_after_sync();
        store_de(prev, (ntiles - 1) & 1);
        }
        __syncthreads();
        if (!issuer) {
        // all MMAs of the encoder block were committed to barA before the last barB commit: complete as well
        // ---- flush the weight-gradient accumulators: TMEM lane = stacked row c, column = h (bias at h = H)
        if (part < 2) {
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
            RowRef r; r.kind = -1; r.idx = 0;
            if (c < NS) r = row_ref(a, tcg, c);
            float* gw = nullptr; float* gb = nullptr;
            if (r.kind == 0) { gw = g.Wp + size_t(r.idx) * H; gb = g.bp + r.idx; }
            else if (r.kind == 1) { gw = a.sv.dG + size_t(r.idx) * H; gb = a.sv.dc + r.idx; }
            else if (r.kind == 2) { gw = g.Wb + size_t(r.idx) * H; gb = g.bb + r.idx; }
            else if (r.kind == 3) { gw = g.Wa + size_t(r.idx) * H; gb = g.ba + r.idx; }
            if (gw) {
                if ((H & 3) == 0) {                           // rows are 16-byte aligned: one vector reduction per 4 columns
#pragma unroll
                    for (int h4 = 0; h4 < 8; ++h4)
                        if (4 * h4 < H) red_add_v4(gw + 4 * h4, w[4 * h4], w[4 * h4 + 1], w[4 * h4 + 2], w[4 * h4 + 3]);
                } else {
#pragma unroll
                    for (int h = 0; h < 32; ++h) if (h < H) atomicAdd(gw + h, w[h]);
                }
#pragma unroll
                for (int h = 0; h < 32; ++h) if (h == H) atomicAdd(gb, w[h]);
            }
        } else if (part == 3) {
            const int j = tid - 3 * 128;
            if (j < H) { atomicAdd(g.wam + j, accal[j]); atomicAdd(g.was + j, accal[32 + j]); }
            if (j == 31) { atomicAdd(g.bam, accal[31]); atomicAdd(g.bas, accal[63]); }
        }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}


// =================================================================================================================
// Forward of the heads on tcgen05 (FVAE_PREC_BF16_TC, M == 128, H <= 31, K <= 32).  One CTA walks whole dates.
//   loop 1 over the 128-stock tiles of the date:  F_enc^T = Wp . E^T  (TMEM lane = portfolio j, column = stock)
//                                                 F_row   = E . [G; Wb; Wa]^T  (lane = stock)
//        group A (thread = portfolio column): online softmax over the stocks -> max, sum, y_p
//        group B (thread = stock): attention scores -> per-head maximum (the NaN/Inf guard rides in it as +inf)
//   posterior (mapping layer), then loop 2:
//        group A (thread = stock): attention weights p = exp(s - max) -> bf16 tile P, per-head sums;
//                                  pooled += P^T . E as a UMMA whose accumulator stays in TMEM for the date
//        group B (thread = stock): decoder (alpha / beta heads, mu_y, sigma_y, sample, squared error, the column sums
//                                  backward needs)
//   prior head, KL, loss.  FactorVAE.prediction: the decoder runs in a third loop, after the prior.
// Same image (t_b1: hi | lo rows [Wp; G; Wb; Wa] with the bias column) as the backward sweep.
constexpr int FNT = 256;
constexpr uint32_t kFColEnc = 0, kFColRow = 128, kFColPool = 224;      // TMEM columns (256 allocated)

struct FwdSmem {
    uint32_t et, b1s, pt, ys, yp, muz, sgz, mupr, sgpr, pooled, ctx, hm, attm, attl, c1a, c1b, red, bad, gbad, bar, slot, total;
};
__host__ __device__ inline FwdSmem fwd_layout(int M, const TcCols& tcg) {
    FwdSmem s; uint32_t p = 0;
    auto take = [&](uint32_t n) { uint32_t r = p; p += (n + 127u) & ~127u; return r; };
    s.et = take(2 * 8 * kTileChunk);
    s.pt = take(16 * kTileChunk);           // P tile (Kp/8 chunks written; the UMMA M block spans 16)
    s.b1s = take(8u * tcg.NS * 16u);
    s.ys = take(128 * 4);
    s.yp = take(uint32_t(M) * 4);
    s.muz = take(32 * 4); s.sgz = take(32 * 4); s.mupr = take(32 * 4); s.sgpr = take(32 * 4);
    s.pooled = take(32 * 32 * 4); s.ctx = take(32 * 32 * 4); s.hm = take(32 * 32 * 4);
    s.attm = take(32 * 4); s.attl = take(32 * 4); s.c1a = take(32 * 4); s.c1b = take(32 * 4);
    s.red = take(32 * 4); s.bad = take(32 * 4); s.gbad = take(32 * 4);
    s.bar = take(16); s.slot = take(16);
    s.total = p;
    return s;
}

template <typename Op>
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane, Op op) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float keep = up ? v[i + o] : v[i];
            const float send = up ? v[i] : v[i + o];
            v[i] = op(keep, __shfl_xor_sync(0xffffffffu, send, o));
        }
    }
    return v[0];
}

__global__ void __launch_bounds__(FNT, 2) heads_tc_fwd_kernel(HeadsArgs a, TcCols tcg) {
    const uint64_t nstep = noise_step(a);
    extern __shared__ __align__(1024) uint8_t smem[];
    const int H = a.H, K = a.K, M = a.M, NS = tcg.NS, Kp = tcg.Kp, Hp8 = tcg.Hp8;
    const int tid = threadIdx.x, row = tid & 127, grp = tid >> 7, warp = tid >> 5, lane = tid & 31;
    const FwdSmem L = fwd_layout(M, tcg);
    uint8_t* Et = smem + L.et;
    uint8_t* Pt = smem + L.pt;
    uint8_t* B1s = smem + L.b1s;
    float* ys = reinterpret_cast<float*>(smem + L.ys);
    float* yp = reinterpret_cast<float*>(smem + L.yp);
    float* muz = reinterpret_cast<float*>(smem + L.muz);
    float* sgz = reinterpret_cast<float*>(smem + L.sgz);
    float* mupr = reinterpret_cast<float*>(smem + L.mupr);
    float* sgpr = reinterpret_cast<float*>(smem + L.sgpr);
    float* pooled = reinterpret_cast<float*>(smem + L.pooled);   // [k][h], row stride H
    float* ctx = reinterpret_cast<float*>(smem + L.ctx);
    float* hm = reinterpret_cast<float*>(smem + L.hm);
    int* attm = reinterpret_cast<int*>(smem + L.attm);           // per-head maximum as float bits (scores are >= 0)
    float* attl = reinterpret_cast<float*>(smem + L.attl);
    float* c1a = reinterpret_cast<float*>(smem + L.c1a);
    float* c1b = reinterpret_cast<float*>(smem + L.c1b);
    float* red = reinterpret_cast<float*>(smem + L.red);
    int* bad = reinterpret_cast<int*>(smem + L.bad);
    int* gbad = reinterpret_cast<int*>(smem + L.gbad);          // collapsed key row of the head is not finite: the guard trips on every date
    uint64_t* barM = reinterpret_cast<uint64_t*>(smem + L.bar);
    uint64_t* barP = barM + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + L.slot);

    if (warp == 0) tmem_alloc<256>(slot);
    if (tid == 0) { mbar_init(barM, 1); mbar_init(barP, 1); mbar_fence_init(); }
    {
        const uint4* s1 = static_cast<const uint4*>(a.sv.t_b1);
        for (int i = tid; i < 8 * NS; i += FNT) reinterpret_cast<uint4*>(B1s)[i] = s1[i];
        for (int i = tid; i < 16 * 128; i += FNT) reinterpret_cast<uint4*>(Pt)[i] = make_uint4(0, 0, 0, 0);
        if (tid < 32) {          // the images hold sanitised rows, so a non-finite G_k / c_k has to be remembered here
            int nb = 0;
            if (tid < K) {
                if (!(fabsf(a.sv.cvec[tid]) <= FLT_MAX)) nb = 1;
                for (int h = 0; h < H; ++h) if (!(fabsf(a.sv.G[size_t(tid) * H + h]) <= FLT_MAX)) nb = 1;
            }
            gbad[tid] = nb;
        }
    }
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *slot;
    const uint32_t lane_base = tmem + (uint32_t((warp & 3) * 32) << 16);
    const uint32_t et_addr = smem_u32(Et), pt_addr = smem_u32(Pt), b1s_addr = smem_u32(B1s);
    const float inv_tau = 1.f / sqrtf(float(H) + 1e-6f);
    const bool train_path = !a.predict;
    uint32_t phM = 0, phP = 0;

    // D[dcol..) = A . B^T with both operands split hi | lo (chunks 0-3 | 4-7, K = 32): hi.hi + lo.hi + hi.lo
    auto issue_split = [&](uint32_t a_addr, uint32_t a_lbo, uint32_t b_addr, uint32_t b_lbo, uint32_t N, uint32_t dcol) {
        const uint32_t idesc = make_idesc_bf16(kTileRows, N, false, false);
        auto mm = [&](uint32_t ac, uint32_t bc, uint32_t acc) {
            mma_bf16_ss(tmem + dcol, make_smem_desc(a_addr + ac * a_lbo, a_lbo, 128), make_smem_desc(b_addr + bc * b_lbo, b_lbo, 128), idesc, acc);
        };
        mm(0, 0, 0); mm(2, 2, 1); mm(4, 0, 1); mm(6, 2, 1); mm(0, 4, 1); mm(2, 6, 1);
    };

    for (int d = blockIdx.x; d < a.B; d += gridDim.x) {
        const int p0 = a.date_ptr[d], n = a.date_ptr[d + 1] - p0;
        if (n <= 0) { if (tid == 0 && a.out.date_loss) a.out.date_loss[d] = nanf(""); continue; }
        const int ntile = (n + 127) / 128;
        const float coefN = 2.f / (float(n) * float(a.B));
        if (tid < 32) { attm[tid] = 0; attl[tid] = 0.f; c1a[tid] = 0.f; c1b[tid] = 0.f; bad[tid] = 0; }
        float m_e = -INFINITY, l_e = 0.f, acc_e = 0.f;          // group A in loop 1: online softmax state of portfolio column `row`
        float rec_part = 0.f;
        __syncthreads();

        auto load_e = [&](int t, float (&v)[16]) {               // my 16 columns [16 grp, 16 grp + 16) of my row (ones column at H)
            const int i = t * 128 + row;
            const bool valid = i < n;
            const float* src = a.e + size_t(p0 + i) * H;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int h = grp * 16 + q;
                v[q] = (valid && h < H) ? src[h] : ((valid && h == H) ? 1.f : 0.f);
            }
        };
        auto store_e = [&](int buf, const float (&v)[16]) {
            uint8_t* eb = Et + buf * 8 * kTileChunk;
            float lo8[8], hi8[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) { lo8[q] = v[q]; hi8[q] = v[8 + q]; }
            split_store8(eb + tile_off(128, row, 2 * grp), eb + tile_off(128, row, 4 + 2 * grp), lo8);
            split_store8(eb + tile_off(128, row, 2 * grp + 1), eb + tile_off(128, row, 4 + 2 * grp + 1), hi8);
        };
        // attention score of (my stock, head k) from the accumulator value: relu(dropout(f / tau)); non-finite -> +inf
        auto att_score = [&](float f, int u, int k) {
            const float x = relu_nan(f * inv_tau * keep_factor(a, nstep, u, k));
            return (fabsf(x) <= FLT_MAX) ? x : INFINITY;
        };

        auto tile_loop = [&](bool do_enc, bool do_attmax, bool do_attp, bool do_dec) {
            float ev[16];
            load_e(0, ev);
            for (int t = 0; t < ntile; ++t) {
                store_e(t & 1, ev);
                if (tid < 128) ys[tid] = (train_path && t * 128 + tid < n) ? a.y[p0 + t * 128 + tid] : 0.f;
                if (do_attp && t > 0) { mbar_wait(barP, phP); phP ^= 1; }          // pooled MMA of the previous tile: P tile free
                fence_async_smem();
                __syncthreads();
                if (tid == 0) {
                    tc_fence_after_sync();
                    const uint32_t ea = et_addr + (t & 1) * 8 * kTileChunk;
                    if (do_enc) issue_split(b1s_addr, uint32_t(NS) * 16u, ea, kTileChunk, 128, kFColEnc);
                    issue_split(ea, kTileChunk, b1s_addr + 128 * 16, uint32_t(NS) * 16u, uint32_t(NS - 128), kFColRow);
                    mma_commit(barM);
                }
                if (t + 1 < ntile) load_e(t + 1, ev);
                mbar_wait(barM, phM); phM ^= 1;
                tc_fence_after_sync();
                const int nv = min(128, n - t * 128);
                const int i = t * 128 + row;
                const bool valid = i < n;
                const int u = p0 + i;
                if (grp == 0) {
                    if (do_enc) {        // thread = portfolio column `row`: columns of my TMEM lane are the stocks of the tile
                        float tmax = -INFINITY;
                        for (int g16 = 0; g16 * 16 < nv; ++g16) {
                            float f[16];
                            tmem_ld16(lane_base + kFColEnc + uint32_t(g16 * 16), f);
#pragma unroll
                            for (int q = 0; q < 16; ++q) tmax = fmaxf(tmax, (g16 * 16 + q < nv) ? f[q] : -INFINITY);
                        }
                        if (tmax > m_e) {
                            const float sc = ex2_fast((m_e - tmax) * kL2E);
                            l_e *= sc; acc_e *= sc; m_e = tmax;
                        }
                        const float mb = -m_e * kL2E;
                        for (int g16 = 0; g16 * 16 < nv; ++g16) {
                            float f[16];
                            tmem_ld16(lane_base + kFColEnc + uint32_t(g16 * 16), f);
#pragma unroll
                            for (int q = 0; q < 16; ++q) {
                                const float pz = (g16 * 16 + q < nv) ? ex2_fast(fmaf(f[q], kL2E, mb)) : 0.f;
                                l_e += pz;
                                acc_e = fmaf(pz, ys[g16 * 16 + q], acc_e);
                            }
                        }
                    }
                    if (do_attp) {       // thread = stock: attention weights (unnormalised) -> P tile, per-head sums
                        float r[32];
#pragma unroll
                        for (int kc = 0; kc < 4; ++kc) {
                            float f[8], pz[8];
                            if (kc < Kp / 8) tmem_ld8(lane_base + kFColRow + uint32_t(kc * 8), f);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const int k = kc * 8 + q;
                                float pv = 0.f;
                                if (kc < Kp / 8 && valid && k < K && !bad[k]) pv = ex2_fast((att_score(f[q], u, k) - __int_as_float(attm[k])) * kL2E);
                                pz[q] = pv; r[k] = pv;
                            }
                            if (kc < Kp / 8) store8(Pt + tile_off(128, row, kc), pz);
                        }
                        const float sl = warp_transpose_reduce(r, lane, [](float x, float y) { return x + y; });
                        atomicAdd(attl + lane, sl);
                    }
                } else {
                    if (do_attmax) {     // thread = stock: per-head maximum of the scores over the stocks
                        float r[32];
#pragma unroll
                        for (int kc = 0; kc < 4; ++kc) {
                            float f[8];
                            if (kc < Kp / 8) tmem_ld8(lane_base + kFColRow + uint32_t(kc * 8), f);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const int k = kc * 8 + q;
                                r[k] = (kc < Kp / 8 && valid && k < K) ? att_score(f[q], u, k) : 0.f;
                            }
                        }
                        const float mx = warp_transpose_reduce(r, lane, [](float x, float y) { return fmaxf(x, y); });
                        atomicMax(attm + lane, __float_as_int(mx));
                    }
                    if (do_dec) {        // thread = stock: decoder (module.py:107-123)
                        float bt[32];
                        float mu = 0.f, var = 0.f;
#pragma unroll
                        for (int kc = 0; kc < 4; ++kc) {
                            float f[8];
                            if (kc < Kp / 8) tmem_ld8(lane_base + kFColRow + uint32_t(Kp + kc * 8), f);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const int k = kc * 8 + q;
                                const float b = (kc < Kp / 8 && k < K) ? f[q] : 0.f;
                                bt[k] = b;
                                if (k < K) { mu = fmaf(b, muz[k], mu); var = fmaf(b * b, sgz[k] * sgz[k], var); }
                            }
                        }
                        float amu = a.w.bam[0], asp = a.w.bas[0];
#pragma unroll
                        for (int jc = 0; jc < 4; ++jc) {
                            float f[8];
                            if (jc < Hp8 / 8) tmem_ld8(lane_base + kFColRow + uint32_t(2 * Kp + jc * 8), f);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const int j = jc * 8 + q;
                                if (jc < Hp8 / 8 && j < H) {
                                    const float ha = lrelu(f[q]);
                                    amu = fmaf(a.w.wam[j], ha, amu);
                                    asp = fmaf(a.w.was[j], ha, asp);
                                }
                            }
                        }
                        float dmy = 0.f, dvv = 0.f;
                        if (valid) {
                            const float asig = softplus(asp);
                            mu += amu;
                            const float sy = sqrtf(var + asig * asig + 1e-6f);
                            const float ep = eps_of(a, nstep, u);
                            const float yh = fmaf(ep, sy, mu);
                            a.out.yhat[u] = yh; a.out.mu_y[u] = mu; a.out.sigma_y[u] = sy;
                            if (train_path) {
                                const float dlt = yh - ys[row];
                                rec_part = fmaf(dlt, dlt, rec_part);
                                dmy = coefN * dlt;
                                dvv = dmy * ep / (2.f * sy);
                            }
                        }
                        if (train_path) {        // column sums for backward: sum_i beta_ik dmu_y_i, sum_i beta_ik^2 dvar_i
                            float r[32];
#pragma unroll
                            for (int k = 0; k < 32; ++k) r[k] = bt[k] * dmy;
                            const float s1 = warp_transpose_reduce(r, lane, [](float x, float y) { return x + y; });
#pragma unroll
                            for (int k = 0; k < 32; ++k) r[k] = bt[k] * bt[k] * dvv;
                            const float s2 = warp_transpose_reduce(r, lane, [](float x, float y) { return x + y; });
                            atomicAdd(c1a + lane, s1);
                            atomicAdd(c1b + lane, s2);
                        }
                    }
                }
                if (do_attp) fence_async_smem();
                tc_fence_before_sync();
                __syncthreads();
                if (do_attp && tid == 0) {
                    tc_fence_after_sync();
                    issue_wgrad_acc(tmem, kFColPool, pt_addr, 0, et_addr + (t & 1) * 8 * kTileChunk, kHK, t > 0);
                    mma_commit(barP);
                }
            }
            if (do_attp) { mbar_wait(barP, phP); phP ^= 1; tc_fence_after_sync(); }
        };

        // ---- loop 1: softmax statistics
        tile_loop(train_path, true, false, false);
        if (grp == 0 && train_path) {
            a.sv.enc_m[size_t(d) * M + row] = m_e;
            a.sv.enc_l[size_t(d) * M + row] = l_e;
            const float v = acc_e / l_e;
            yp[row] = v;
            a.sv.yp[size_t(d) * M + row] = v;
        }
        if (tid < 32) {
            const float mk = __int_as_float(attm[tid]);
            const int bd = (tid < K && (gbad[tid] || !(fabsf(mk) <= FLT_MAX))) ? 1 : 0;
            bad[tid] = bd;
            if (bd) attm[tid] = 0;
        }
        __syncthreads();
        // ---- posterior (module.py:48-49) and the :117 clamp: 8 threads per factor
        if (train_path) {
            const int k = tid >> 3, part = tid & 7;
            float mu = 0.f, pre = 0.f;
            if (k < K) {
                const float* wm = a.w.Wmu + size_t(k) * M;
                const float* ws = a.w.Wsig + size_t(k) * M;
                for (int j = part; j < M; j += 8) { mu = fmaf(wm[j], yp[j], mu); pre = fmaf(ws[j], yp[j], pre); }
            }
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) { mu += __shfl_xor_sync(0xffffffffu, mu, o); pre += __shfl_xor_sync(0xffffffffu, pre, o); }
            if (k < K && part == 0) {
                mu += a.w.bmu[k]; pre += a.w.bsig[k];
                float sg = softplus(pre);
                const int cl = (sg == 0.f);
                if (cl) sg = kSigmaFloor;
                muz[k] = mu; sgz[k] = sg;
                a.out.mu_post[size_t(d) * K + k] = mu;
                a.out.sigma_post[size_t(d) * K + k] = sg;
                a.sv.pre_sg_post[size_t(d) * K + k] = pre;
                a.sv.clamp_post[size_t(d) * K + k] = cl;
            }
        }
        __syncthreads();
        // ---- loop 2: attention weights + pooled hidden state (and the decoder when the posterior feeds it)
        tile_loop(false, false, true, train_path);
        if (warp == 0) {                                         // pooled: TMEM lane = head k, column = h
            float v[32];
            {
                float t8[8];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    tmem_ld8(lane_base + kFColPool + 8u * q4, t8);
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q4 * 8 + q] = t8[q];
                }
            }
            const int k = lane;
            if (k < K) {
                const float l = attl[k];
                const bool bd = bad[k] != 0;
                a.sv.att_m[size_t(d) * K + k] = __int_as_float(attm[k]);
                a.sv.att_l[size_t(d) * K + k] = l;
                a.sv.bad[size_t(d) * K + k] = bd ? 1 : 0;
                const float inv = bd ? 0.f : 1.f / l;
#pragma unroll
                for (int h = 0; h < 32; ++h)
                    if (h < H) {
                        const float pv = bd ? 0.f : v[h] * inv;
                        pooled[k * H + h] = pv;
                        a.sv.pooled[(size_t(d) * K + k) * H + h] = pv;
                    }
            }
        }
        tc_fence_before_sync();
        __syncthreads();
        // ---- prior: ctx_k = Wv_k pooled_k + bv_k (zeros if the guard tripped), shared MLP head (module.py:169-188)
        for (int idx = tid; idx < K * H; idx += FNT) {
            const int k = idx / H, j = idx % H;
            float v = 0.f;
            if (!bad[k]) {
                const float* wv = a.w.Wv + (size_t(k) * H + j) * H;
                v = a.w.bv[size_t(k) * H + j];
                for (int h = 0; h < H; ++h) v = fmaf(wv[h], pooled[k * H + h], v);
            }
            ctx[idx] = v;
            a.sv.ctx[size_t(d) * K * H + idx] = v;
        }
        __syncthreads();
        for (int idx = tid; idx < K * H; idx += FNT) {
            const int k = idx / H, j = idx % H;
            float v = a.w.bl[j];
            const float* wl = a.w.Wl + size_t(j) * H;
            for (int h = 0; h < H; ++h) v = fmaf(wl[h], ctx[k * H + h], v);
            a.sv.hm_pre[size_t(d) * K * H + idx] = v;
            hm[idx] = lrelu(v);
        }
        __syncthreads();
        for (int k = tid; k < K; k += FNT) {
            float mu = a.w.bpm[0], pre = a.w.bps[0];
            for (int j = 0; j < H; ++j) { mu = fmaf(a.w.wpm[j], hm[k * H + j], mu); pre = fmaf(a.w.wps[j], hm[k * H + j], pre); }
            float sg = softplus(pre);
            const int cl = (sg == 0.f);
            if (cl) sg = kSigmaFloor;                          // module.py:264-265 (and :117 in prediction)
            mupr[k] = mu; sgpr[k] = sg;
            a.out.mu_prior[size_t(d) * K + k] = mu;
            a.out.sigma_prior[size_t(d) * K + k] = sg;
            a.sv.pre_sg_prior[size_t(d) * K + k] = pre;
            a.sv.clamp_prior[size_t(d) * K + k] = cl;
            if (a.predict) { muz[k] = mu; sgz[k] = sg; }
        }
        __syncthreads();
        if (a.predict) {
            tile_loop(false, false, false, true);              // decoder fed by the prior (module.py:273-278)
            __syncthreads();
            continue;
        }
        if (tid < K) { a.sv.c1_mu[size_t(d) * K + tid] = c1a[tid]; a.sv.c1_sg[size_t(d) * K + tid] = c1b[tid]; }
        const float rec = block_sum(rec_part, red) / float(n);   // F.mse_loss: mean over stocks
        float klp = 0.f;
        for (int k = tid; k < K; k += FNT) {                     // module.py:247
            const float m1 = muz[k], s1 = sgz[k], m2 = mupr[k], s2 = sgpr[k];
            klp += logf(s2 / s1) + (s1 * s1 + (m1 - m2) * (m1 - m2)) / (2.f * s2 * s2) - 0.5f;
        }
        const float kl = block_sum(klp, red);
        if (tid == 0) a.out.date_loss[d] = rec + kl;
        __syncthreads();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace

int64_t heads_tc_image_bytes(int H, int K, int M, int which) {
    const TcCols c = tc_cols(H, K, M);
    return which == 1 ? int64_t(8) * c.NS * 16 : int64_t(c.NS) * 64;
}

int heads_tc_prep(const HeadsArgs& a, cudaStream_t stream) {
    const TcCols c = tc_cols(a.H, a.K, a.M);
    heads_tc_prep_kernel<<<(c.NS * kHK + 255) / 256, 256, 0, stream>>>(a, c); count_launch();
    return int(cudaGetLastError());
}

int heads_tc_forward(const HeadsArgs& a, cudaStream_t stream) {
    const TcCols c = tc_cols(a.H, a.K, a.M);
    const FwdSmem L = fwd_layout(a.M, c);
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    cudaError_t e = cudaFuncSetAttribute(heads_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(L.total));
    if (e != cudaSuccess) return int(e);
    const int grid = a.B < 2 * sms ? a.B : 2 * sms;
    heads_tc_fwd_kernel<<<grid, FNT, L.total, stream>>>(a, c); count_launch();
    return int(cudaGetLastError());
}

int heads_tc_sweep(const HeadsArgs& a, const HeadsG& g, float* dE, cudaStream_t stream) {
    const TcCols c = tc_cols(a.H, a.K, a.M);
    const SweepSmem L = sweep_layout(a.M, c);
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    cudaError_t e = cudaFuncSetAttribute(heads_tc_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(L.total));
    if (e != cudaSuccess) return int(e);
    heads_tc_sweep_kernel<<<sms, TNT_ALL, L.total, stream>>>(a, g, dE, c); count_launch();
    return int(cudaGetLastError());
}

}  // namespace fvae
