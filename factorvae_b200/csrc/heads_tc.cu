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
                    const float coefN = 2.f / (float(cur.n) * float(a.B));
                    v1 = coefN * (a.out.yhat[u] - a.y[u]);
                    v2 = v1 * eps_of(a, u) / (2.f * a.out.sigma_y[u]);
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

int heads_tc_sweep(const HeadsArgs& a, const HeadsG& g, float* dE, cudaStream_t stream) {
    const TcCols c = tc_cols(a.H, a.K, a.M);
    const SweepSmem L = sweep_layout(a.M, c);
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    cudaError_t e = cudaFuncSetAttribute(heads_tc_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(L.total));
    if (e != cudaSuccess) return int(e);
    int rc = heads_tc_prep(a, stream);
    if (rc != 0) return rc;
    heads_tc_sweep_kernel<<<sms, TNT_ALL, L.total, stream>>>(a, g, dE, c); count_launch();
    return int(cudaGetLastError());
}

}  // namespace fvae
