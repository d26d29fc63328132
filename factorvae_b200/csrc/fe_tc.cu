// FeatureExtractor, FVAE_PREC_BF16_TC: the whole FeatureExtractor forward + backward on tcgen05 tensor
// cores (bf16 operands, fp32 accumulation in TMEM), sm_100a.  Restates reference module.py:22-31 and its
// autograd (train_model.py:29); the fp32 CUDA-core restatement in fe_f32.cu is the GPU-side cross-check.
//
// Work unit: an ITEM = (sequence tile st of 128 stocks, time step t) = 128 rows of the panel = one UMMA M.
//
//   K1 front_fwd  per item : x rows -> LayerNorm (fp32) -> xhat bf16 tile -> MMA [128x160]=(xhat . W1g^T)
//                            -> +b1f, LeakyReLU, bf16 tile -> MMA [128xNC]=(u . W_ih^T) -> + bias -> GI tile
//                            (LayerNorm's affine is folded into the weights: W1g = W1 diag(gamma),
//                             b1f = b1 + W1 beta; the gate columns are permuted into blocks of 8 units
//                             [r8|z8|n8] so one hidden unit's three gates sit in adjacent 16-byte chunks)
//   K2 gru_fwd    per tile : for t: MMA [128xNC] = (h_{t-1} . W_hh^T) -> gates (fp32) -> h_t (fp32 regs,
//                            bf16 operand tile in smem, bf16 HALL tile in HBM for backward)
//   K3 gru_bwd    per tile : BPTT; per step MMA gh (recompute), gate gradients (fp32) -> dGI tile (in place)
//                            and dgh tile, MMA dh += dgh . W_hh, MMA dW_hh += dgh^T . [h_{t-1} | 1] (TMEM
//                            accumulators live across all tiles of the CTA; one flush per CTA)
//   K1 also saves, per item, the xhat and u operand tiles (bf16) and the LeakyReLU' sign bits, so backward never
//   touches the panel again and never recomputes GEMM 1:
//   K4a q_from_tiles per item: du = dGI . W_ih (MMA), dpre = du * LeakyReLU' (epilogue), Q += dpre^T . [xhat | 1]
//   K4b wih_from_u   per item: dWih += dGI^T . [u | 1]
//                            (both stream their operand tiles through a cp.async ring; the weight-gradient MMAs read
//                             the SAME smem tiles as MN-major operands; accumulators live in TMEM for the whole kernel)
//   K5 post                : dW1 = Q diag(gamma) + db1 beta^T, dgamma = sum_o W1 .* Q, dbeta = W1^T db1,
//                            un-permute dW_ih / db_ih.   (x is data: no LayerNorm input gradient, so the
//                            dxn = dpre . W1 contraction of the reference's autograd is never computed)
//
// All operand tiles use the chunk-major SWIZZLE_NONE layout of tc_sm100.cuh.
#include <cstdlib>

#include "fe.cuh"
#include "tc_sm100.cuh"

namespace fvae {
namespace {

using namespace tc;

constexpr int CP = 160;          // C padded (K of the row GEMMs, N of GEMM 1); column C holds the constant 1
constexpr int KCH = CP / 8;      // 20 chunks of 8 features
constexpr int TM = 128;          // rows per item = UMMA M = TMEM lanes
constexpr uint32_t TILE_CH = TM * 16;            // bytes of one 8-column chunk of a 128-row tile (2048)
constexpr uint32_t A_BYTES = KCH * TILE_CH;      // 40960: a [128 x 160] bf16 tile
constexpr uint32_t W1_BYTES = KCH * CP * 16;     // 51200: a [160 x 160] bf16 image

__host__ __device__ inline int pad16(int v) { return (v + 15) & ~15; }
__host__ __device__ inline int nb8_of(int H) { return (H + 7) / 8; }
// Gate columns are permuted into blocks of 8 hidden units, [r8 | z8 | n8] = three 16-byte chunks per block, so a thread
// finds the three gates of its units in adjacent chunks.  When the last block holds at most 4 units it is COMPACT:
// [r4 z4 | n4 0000] = two chunks (H = 20: 24 + 24 + 16 = 64 columns instead of 80 -> 20 % less GI / dGI traffic and a
// 64-column TMEM accumulator for the forward GRU).
// ... used only where it actually shrinks the padded column count (H = 20: 64 < 80; H = 60: 192 either way -> plain blocks)
__host__ __device__ inline bool gate_compact(int H) {
    const int r = H & 7, nb = nb8_of(H);
    return r >= 1 && r <= 4 && pad16(24 * (nb - 1) + 16) < pad16(24 * nb);
}
__host__ __device__ inline int nc_of(int H) {                                     // gate columns (permuted, padded)
    return gate_compact(H) ? pad16(24 * (nb8_of(H) - 1) + 16) : pad16(24 * nb8_of(H));
}
__host__ __device__ inline int hp_of(int H) { return pad16(H + 1); }              // hidden columns + the ones column
// gate g = gate*H + j  ->  column of the permuted layout
__host__ __device__ inline int perm_col(int gate, int j, int H) {
    const int blk = j >> 3, u = j & 7;
    if (gate_compact(H) && blk == nb8_of(H) - 1) return blk * 24 + (gate < 2 ? gate * 4 + u : 8 + u);
    return blk * 24 + gate * 8 + u;
}
// column -> (gate, j); returns false for padding columns
__host__ __device__ inline bool unperm_col(int col, int H, int& gate, int& j) {
    const int blk = col / 24, rem = col % 24;
    if (gate_compact(H) && blk == nb8_of(H) - 1) {
        if (rem >= 16) return false;
        gate = rem < 8 ? (rem >> 2) : 2;
        const int u = rem < 8 ? (rem & 3) : rem - 8;
        j = blk * 8 + u;
        return u < 4 && j < H;
    }
    gate = rem >> 3;
    j = blk * 8 + (rem & 7);
    return blk < nb8_of(H) && j < H;
}

// ---- workspace -----------------------------------------------------------------------------------
struct TcWs {
    __nv_bfloat16 *w1g, *wih, *wihT, *whh, *whhT;   // operand images
    __nv_bfloat16 *w1n;                             // W1 diag(gamma) image WITHOUT the bias column (TMA kernels: bias added in the epilogue)
    float *b1f, *bgi, *bhn;                         // [CP], [NC], [HP]
    float *w1s;                                     // [CP] row sums of the bf16-rounded W1 diag(gamma) image (LayerNorm-after-GEMM fold)
    __nv_bfloat16 *gi;      // [NT][T][NC/8][128][8]   gate pre-activations, then (backward) their gradients
    __nv_bfloat16 *hall;    // [NT][T][HP/8][128][8]   h_t operand tiles (column H = 1)
    __nv_bfloat16 *xh;      // [NT][T][CP/8][128][8]   xhat = LayerNorm(x) operand tiles (column C = 1), saved for backward
    __nv_bfloat16 *u;       // [NT][T][CP/8][128][8]   u = LeakyReLU(pre) operand tiles (column C = 1), saved for backward
    unsigned long long *mask;  // [NT][T][4][128]      LeakyReLU' sign bits of my 40 columns (bit b: pre > 0)
    float2 *stats;             // [NT][T][128]         row statistics (-mean rstd, rstd) saved by the TMA front kernel
    float *q;               // [2*128][CP]  Q = dpre^T [xhat|1]
    float *dwih;            // [2*128][CP]  dGI^T [u|1]   (permuted rows)
    int64_t bytes;
};

TcWs carve_tc(const FeDims& d, void* base) {
    TcWs w;
    char* p = static_cast<char*>(base);
    auto take = [&](int64_t n) { char* r = p; p += (n + 255) / 256 * 256; return r; };
    const int NC = nc_of(d.H), HP = hp_of(d.H);
    const int64_t NT = (int64_t(d.S) + TM - 1) / TM;
    w.w1g = reinterpret_cast<__nv_bfloat16*>(take(W1_BYTES));
    w.wih = reinterpret_cast<__nv_bfloat16*>(take(int64_t(KCH) * NC * 16));
    w.wihT = reinterpret_cast<__nv_bfloat16*>(take(int64_t(NC / 8) * CP * 16));
    w.whh = reinterpret_cast<__nv_bfloat16*>(take(int64_t(HP / 8) * NC * 16));
    w.whhT = reinterpret_cast<__nv_bfloat16*>(take(int64_t(NC / 8) * HP * 16));
    w.w1n = reinterpret_cast<__nv_bfloat16*>(take(W1_BYTES));
    w.w1s = reinterpret_cast<float*>(take(CP * 4));
    w.b1f = reinterpret_cast<float*>(take(CP * 4));
    w.bgi = reinterpret_cast<float*>(take(NC * 4));
    w.bhn = reinterpret_cast<float*>(take(HP * 4));
    w.gi = reinterpret_cast<__nv_bfloat16*>(take(NT * d.T * int64_t(NC / 8) * TILE_CH));
    w.hall = reinterpret_cast<__nv_bfloat16*>(take(NT * d.T * int64_t(HP / 8) * TILE_CH));
    w.xh = reinterpret_cast<__nv_bfloat16*>(take(NT * d.T * int64_t(A_BYTES)));
    // u tiles are saved only when backward cannot rebuild them (tc_wih_recompute_kernel covers NC <= 128)
    w.u = NC <= 128 ? nullptr : reinterpret_cast<__nv_bfloat16*>(take(NT * d.T * int64_t(A_BYTES)));
    w.mask = reinterpret_cast<unsigned long long*>(take(NT * d.T * int64_t(4 * TM * 8)));
    w.stats = reinterpret_cast<float2*>(take(NT * d.T * int64_t(TM * 8)));
    w.q = reinterpret_cast<float*>(take(int64_t(256) * CP * 4));
    w.dwih = reinterpret_cast<float*>(take(int64_t(256) * CP * 4));
    w.bytes = p - static_cast<char*>(base);
    return w;
}

// ---- K0: operand images (once per step; weights change every step) ------------------------------------
struct PrepArgs {
    int C, H, NC, HP;
    const float *ln_w, *ln_b, *W1, *b1, *Wih, *Whh, *bih, *bhh;
    TcWs ws;
};

__device__ __forceinline__ void put_img(__nv_bfloat16* img, int rows, int row, int k, float v) {
    img[(size_t(k >> 3) * rows + row) * 8 + (k & 7)] = __float2bfloat16(v);
}

__global__ void tc_prep_kernel(PrepArgs a) {
    const int C = a.C, H = a.H, NC = a.NC, HP = a.HP;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    // W1g[n][k] = W1[n][k] * gamma[k]
    for (int idx = tid; idx < CP * CP; idx += nth) {
        const int n = idx / CP, k = idx % CP;
        const float wv = (n < C && k < C) ? a.W1[n * C + k] * a.ln_w[k] : 0.f;
        if (k != C && k != C + 1) put_img(a.ws.w1n, CP, n, k, wv);     // rows C, C+1 of the TMA image: the LayerNorm fold (written below)
        if (k == C) continue;                    // column C carries the folded bias (written below by another thread)
        put_img(a.ws.w1g, CP, n, k, wv);
    }
    // W_ih (rows permuted) and its transpose
    for (int idx = tid; idx < NC * CP; idx += nth) {
        const int col = idx / CP, k = idx % CP;
        int gate, j;
        const bool ok = unperm_col(col, H, gate, j) && k < C;
        const float v = ok ? a.Wih[(gate * H + j) * C + k] : 0.f;
        if (k != C) put_img(a.ws.wih, NC, col, k, v);       // B[n=col][k]; column C carries the folded bias
        put_img(a.ws.wihT, CP, k, col, v);      // B[n=k(feature)][k=col]
    }
    // W_hh (rows permuted) and its transpose
    for (int idx = tid; idx < NC * HP; idx += nth) {
        const int col = idx / HP, k = idx % HP;
        int gate, j;
        const bool ok = unperm_col(col, H, gate, j) && k < H;
        const float v = ok ? a.Whh[(gate * H + j) * H + k] : 0.f;
        put_img(a.ws.whh, NC, col, k, v);
        put_img(a.ws.whhT, HP, k, col, v);
    }
    // biases ride in column C of the images (column C of the xhat / u operand tiles is the constant 1)
    for (int w0 = tid; w0 < CP * 8; w0 += nth) {             // 8 threads per output (CP * 8 is a multiple of 32)
        const int n = w0 >> 3, part = w0 & 7;
        float v = 0.f;
        if (n < C)
            for (int k = part; k < C; k += 8) v = fmaf(a.W1[n * C + k], a.ln_b[k], v);
#pragma unroll
        for (int s = 4; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (part == 0) {
            if (n < C) v += a.b1[n];
            a.ws.b1f[n] = v;
            put_img(a.ws.w1g, CP, n, C, v);
        }
        // w1s[n] = sum_k bf16(W1[n][k] gamma[k]): the sum of the ROUNDED operands the tensor core multiplies
        float ws_ = 0.f;
        if (n < C)
            for (int k = part; k < C; k += 8) ws_ += __bfloat162float(__float2bfloat16(a.W1[n * C + k] * a.ln_w[k]));
#pragma unroll
        for (int s = 4; s > 0; s >>= 1) ws_ += __shfl_xor_sync(0xffffffffu, ws_, s);
        if (part == 0) {
            a.ws.w1s[n] = ws_;
            // TMA kernels: the x stage carries (1/rstd, mean) in columns C, C+1, so acc = x.W1g + b1f / rstd - mean w1s = pre / rstd
            put_img(a.ws.w1n, CP, n, C, n < C ? v : 0.f);
            put_img(a.ws.w1n, CP, n, C + 1, n < C ? -ws_ : 0.f);
        }
    }
    for (int col = tid; col < NC; col += nth) {
        int gate, j;
        float v = 0.f;
        if (unperm_col(col, H, gate, j)) v = a.bih[gate * H + j] + (gate < 2 ? a.bhh[gate * H + j] : 0.f);
        a.ws.bgi[col] = v;
        put_img(a.ws.wih, NC, col, C, v);
    }
    for (int j = tid; j < HP; j += nth) a.ws.bhn[j] = j < H ? a.bhh[2 * H + j] : 0.f;
}

#include "fe_tc_front.cuh"   // ItemArgs, staging + LayerNorm, K1 front forward, K4 front backward
#include "fe_tc_tma.cuh"     // warp-specialised TMA-fed front kernels (dense bf16 panels)
#include "fe_tc_split.cuh"   // front backward as two roles in one launch

// ---- K2: GRU forward -------------------------------------------------------------------------------------------
struct GruArgs {
    int S, T, H, NC, HP; int64_t NT;
    TcWs ws;
    float* e;             // [S][H]
    const float* dE;      // [S][H]   (backward)
    float *gWhh, *gbhh;   // gradient sections (backward)
    int gi_ring;          // forward: stream the gate pre-activation tiles through a 2-stage bulk-copy ring in shared memory
};

// Threads: (row of the tile, group of BPT 8-unit blocks) -- 128 * NB8 / BPT threads per CTA.  A thread owns 8*BPT hidden
// units of one row: their gate columns of the accumulator, their chunks of the GI tile and of the h operand tile.
// Each CTA is one latency-bound chain (barrier -> UMMA -> gate math per time step), so throughput comes from resident
// CTAs: small H runs thread-per-row (BPT = NB8, 4 CTAs/SM); large H splits the row (BPT = 1) to shorten the chain.
__host__ __device__ constexpr int gru_bpt(int NB8) { return NB8 <= 3 ? NB8 : 1; }
__host__ __device__ constexpr int gru_threads(int NB8) { return TM * (NB8 / gru_bpt(NB8)); }
__host__ __device__ // compact last block (gate_compact): chunks [r4 z4] [n4 0000] -> the three full-format arrays (units 4..7 are padding)
__device__ __forceinline__ void expand_compact(const float (&a)[8], const float (&b)[8], float (&r)[8], float (&z)[8], float (&n)[8]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) { r[u] = a[u]; z[u] = a[4 + u]; n[u] = b[u]; r[4 + u] = 0.f; z[4 + u] = 0.f; n[4 + u] = 0.f; }
}
constexpr int gru_min_ctas(int NB8) { return NB8 <= 3 ? 4 : (NB8 == 4 ? 2 : 1); }

// CMP: the last gate block is compact (gate_compact(H)); a template parameter so the plain-layout kernels carry none of its code
template <int NB8, uint32_t TCOLS, bool CMP>
__global__ void __launch_bounds__(gru_threads(NB8), (TCOLS == 64 && NB8 <= 3) ? 5 : gru_min_ctas(NB8)) tc_gru_fwd_kernel(GruArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NTHR = gru_threads(NB8), BPT = gru_bpt(NB8);
    const int tid = threadIdx.x, warp = tid >> 5, row = tid & (TM - 1), blk0 = (tid >> 7) * BPT;
    const int H = a.H, NC = a.NC, HP = a.HP, NCH = NC / 8, HCH = HP / 8;
    unsigned char* sWhh = smem;                                   // [HCH][NC][16]
    unsigned char* sH = sWhh + uint32_t(HCH) * NC * 16;           // [HCH][128][16]
    float* sBhn = reinterpret_cast<float*>(sH + uint32_t(HCH) * TILE_CH);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sBhn + HP);
    uint64_t* full = bar + 1;                                     // [2] ring stages filled (bulk copy complete_tx)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 3);
    unsigned char* sGi = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 127) & ~uintptr_t(127));
    const uint32_t gi_bytes = uint32_t(NCH) * TILE_CH;            // one step's tile [NCH][128][16]
    const bool ring = a.gi_ring != 0;
    copy_image(sWhh, a.ws.whh, uint32_t(HCH) * NC * 16);
    for (int i = tid; i < HP; i += NTHR) sBhn[i] = a.ws.bhn[i];
    if (tid == 0) { mbar_init(bar, 1); mbar_init(&full[0], 1); mbar_init(&full[1], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<TCOLS>(tmem_slot);
    // flattened (tile, step) sequence of this CTA: q -> tile blockIdx.x + (q / T) * gridDim.x, step q % T
    const int64_t my_tiles = a.NT > int64_t(blockIdx.x) ? (a.NT - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t total_q = my_tiles * a.T;
    auto fill = [&](int64_t q) {                                  // one thread: bulk copy of step q's tile into stage q & 1
        const int64_t tile = int64_t(blockIdx.x) + (q / a.T) * gridDim.x;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.ws.gi) + size_t(tile * a.T + q % a.T) * gi_bytes;
        mbar_expect_tx(&full[q & 1], gi_bytes);
        bulk_g2s(sGi + (q & 1) * gi_bytes, src, gi_bytes, &full[q & 1]);
    };
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = uint32_t(warp & 3) * 32u;
    uint32_t phase = 0;
    constexpr bool cmp = CMP;                                      // the last block is [r4 z4 | n4]: two chunks
    const int nq = (cmp && blk0 + BPT == NB8) ? 3 * BPT - 1 : 3 * BPT;   // chunks of the GI tile that are mine
    if (ring && tid == 0) { if (total_q > 0) fill(0); if (total_q > 1) fill(1); }
    int64_t q = 0;
    for (int64_t st = blockIdx.x; st < a.NT; st += gridDim.x) {
        float h[8 * BPT];
#pragma unroll
        for (int u = 0; u < 8 * BPT; ++u) h[u] = 0.f;
        const int64_t s = st * TM + row;
        for (int t = 0; t < a.T; ++t, ++q) {
            if (t > 0) {
                fence_async_smem();
                tc_fence_before_sync();
                __syncthreads();
                if (warp == 0) {                              // converged issue: one elected lane per UMMA (tc_sm100.cuh: mma_bf16_ss_w)
                    tc_fence_after_sync();
                    issue_row_gemm_w(tmem, 0, smem_u32(sH), smem_u32(sWhh), NC, NC, HP / 16);
                    mma_commit_w(bar);
                }
            }
            // a CTA barrier (this step's, or the one that closed the previous tile) separates everybody's reads of
            // step q-1's stage from its refill with step q+1
            if (ring && tid == 0 && q >= 1 && q + 1 < total_q) fill(q + 1);
            // my chunks (r | z | n per 8-unit block; [r4 z4 | n4] for a compact last block) of this step's gate pre-activations
            uint4 gq[3 * BPT];
#pragma unroll
            for (int c = 0; c < 3 * BPT; ++c) gq[c] = make_uint4(0, 0, 0, 0);
            if (ring) {
                mbar_wait(&full[q & 1], uint32_t(q >> 1) & 1u);
                const unsigned char* gin = sGi + (q & 1) * gi_bytes;
#pragma unroll
                for (int c = 0; c < 3 * BPT; ++c) if (c < nq) gq[c] = *reinterpret_cast<const uint4*>(gin + tile_off(TM, row, 3 * blk0 + c));
            } else {                                              // direct global loads, in flight while the MMA runs
                const unsigned char* gin = reinterpret_cast<const unsigned char*>(a.ws.gi) + size_t(st * a.T + t) * NCH * TILE_CH;
#pragma unroll
                for (int c = 0; c < 3 * BPT; ++c) if (c < nq) gq[c] = *reinterpret_cast<const uint4*>(gin + tile_off(TM, row, 3 * blk0 + c));
            }
            if (t > 0) {
                mbar_wait(bar, phase);
                phase ^= 1;
                tc_fence_after_sync();
            }
            unsigned char* hout = reinterpret_cast<unsigned char*>(a.ws.hall) + size_t(st * a.T + t) * HCH * TILE_CH;
#pragma unroll
            for (int bb = 0; bb < BPT; ++bb) {
                const int blk = blk0 + bb;
                const bool cblk = cmp && blk == NB8 - 1;
                float ghr[8], ghz[8], ghn[8], gir[8], giz[8], gin8[8];
                if (t > 0) {
                    if (cblk) {
                        float a8[8], b8[8];
                        tmem_ld8x2(tmem_addr(tmem, lane_base, blk * 24), tmem_addr(tmem, lane_base, blk * 24 + 8), a8, b8);
                        expand_compact(a8, b8, ghr, ghz, ghn);
                    } else {
                        tmem_ld8x3(tmem_addr(tmem, lane_base, blk * 24), tmem_addr(tmem, lane_base, blk * 24 + 8),
                                   tmem_addr(tmem, lane_base, blk * 24 + 16), ghr, ghz, ghn);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) ghr[u] = ghz[u] = ghn[u] = 0.f;
                }
                if (cblk) {
                    float a8[8], b8[8];
                    unpack8(gq[3 * bb], a8);
                    unpack8(gq[3 * bb + 1], b8);
                    expand_compact(a8, b8, gir, giz, gin8);
                } else {
                    unpack8(gq[3 * bb], gir);
                    unpack8(gq[3 * bb + 1], giz);
                    unpack8(gq[3 * bb + 2], gin8);
                }
                float hv[8];
#pragma unroll
                for (int p2 = 0; p2 < 4; ++p2) {                 // two hidden units per packed instruction (tc_sm100.cuh: fma2)
                    const int u = 2 * p2, j = blk * 8 + u;
                    const float2 r = sigmoid2(add2(make_float2(gir[u], gir[u + 1]), make_float2(ghr[u], ghr[u + 1])));
                    const float2 z = sigmoid2(add2(make_float2(giz[u], giz[u + 1]), make_float2(ghz[u], ghz[u + 1])));
                    const float2 hb = add2(make_float2(ghn[u], ghn[u + 1]), *reinterpret_cast<const float2*>(sBhn + j));
                    const float2 n = tanh2(fma2(r, hb, make_float2(gin8[u], gin8[u + 1])));
                    const float2 hp = make_float2(h[bb * 8 + u], h[bb * 8 + u + 1]);
                    float2 hn = fma2(z, fma2(n, splat2(-1.f), hp), n);   // (1-z) n + z h
                    if (j >= H) hn.x = 0.f;
                    if (j + 1 >= H) hn.y = 0.f;
                    h[bb * 8 + u] = hn.x; h[bb * 8 + u + 1] = hn.y;
                    hv[u] = (j == H) ? 1.f : hn.x;               // the ones column of the operand tile
                    hv[u + 1] = (j + 1 == H) ? 1.f : hn.y;
                }
                const uint4 pk = make_uint4(pack_bf16(hv[0], hv[1]), pack_bf16(hv[2], hv[3]), pack_bf16(hv[4], hv[5]), pack_bf16(hv[6], hv[7]));
                *reinterpret_cast<uint4*>(sH + tile_off(TM, row, blk)) = pk;
                *reinterpret_cast<uint4*>(hout + tile_off(TM, row, blk)) = pk;
            }
            if (blk0 == 0) {
                for (int b = NB8; b < HCH; ++b) {                // padding chunks (hold the ones column when H % 8 == 0)
                    const uint32_t one = (H >= b * 8 && H < b * 8 + 8) ? (0x3F80u << (16 * (H & 1))) : 0u;
                    uint4 pk = make_uint4(0, 0, 0, 0);
                    const int w = (H - b * 8) >> 1;
                    if (one) { if (w == 0) pk.x = one; else if (w == 1) pk.y = one; else if (w == 2) pk.z = one; else pk.w = one; }
                    *reinterpret_cast<uint4*>(sH + tile_off(TM, row, b)) = pk;
                    *reinterpret_cast<uint4*>(hout + tile_off(TM, row, b)) = pk;
                }
            }
            tc_fence_before_sync();
        }
        if (s < a.S) {
#pragma unroll
            for (int u = 0; u < 8 * BPT; ++u) { const int j = blk0 * 8 + u; if (j < H) a.e[s * H + j] = h[u]; }
        }
        __syncthreads();      // sH is rewritten by the next tile's first step only after everyone is done
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TCOLS>(tmem);
}

// ---- K3: GRU backward (BPTT) -------------------------------------------------------------------------------------
template <int NB8, uint32_t TCOLS, bool CMP>
__global__ void __launch_bounds__(gru_threads(NB8), gru_min_ctas(NB8)) tc_gru_bwd_kernel(GruArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NTHR = gru_threads(NB8), BPT = gru_bpt(NB8), NTB = NB8 / BPT;
    const int tid = threadIdx.x, warp = tid >> 5, row = tid & (TM - 1), tblk = tid >> 7, blk0 = tblk * BPT;   // thread = (row, BPT blocks)
    const int H = a.H, NC = a.NC, HP = a.HP, NCH = NC / 8, HCH = HP / 8;
    const int MB = NC > 128 ? 2 : 1;                              // M blocks of the dW_hh accumulator
    unsigned char* sWhh = smem;                                   // [HCH][NC][16]   B of gh = h W_hh^T
    unsigned char* sWhhT = sWhh + uint32_t(HCH) * NC * 16;        // [NCH][HP][16]   B of dh += dgh W_hh
    unsigned char* sHp = sWhhT + uint32_t(NCH) * HP * 16;         // [HCH][128][16]  h_{t-1} (column H = 1)
    unsigned char* sDgh = sHp + uint32_t(HCH) * TILE_CH;          // [16*MB][128][16] dgh (only NCH chunks are written)
    float* sBhn = reinterpret_cast<float*>(sDgh + uint32_t(16 * MB) * TILE_CH);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBhn + HP);      // 0: gh, 1: dh, 2: dW
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
    // the dh accumulator aliases the first HP columns of the gh accumulator (gh is dead once the gate epilogue has run)
    const uint32_t COL_DH = 0u, COL_DW = uint32_t(NC);
    copy_image(sWhh, a.ws.whh, uint32_t(HCH) * NC * 16);
    copy_image(sWhhT, a.ws.whhT, uint32_t(NCH) * HP * 16);
    for (uint32_t i = tid; i < uint32_t(16 * MB) * TILE_CH / 16; i += NTHR) reinterpret_cast<uint4*>(sDgh)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < HP; i += NTHR) sBhn[i] = a.ws.bhn[i];
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<TCOLS>(tmem_slot);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = uint32_t(warp & 3) * 32u;
    uint32_t ph0 = 0, ph1 = 0, ph2 = 0;
    bool dw_pending = false, dw_started = false;
    constexpr bool cmp = CMP;                                      // the last block is [r4 z4 | n4]: two chunks
    const int nq = (cmp && blk0 + BPT == NB8) ? 3 * BPT - 1 : 3 * BPT;
    // constant chunk of the h operand tile: zeros, with the ones column where H falls into chunk b
    auto const_chunk = [&](int b) {
        uint4 pk = make_uint4(0, 0, 0, 0);
        if (H >= b * 8 && H < b * 8 + 8) {
            const uint32_t one = 0x3F80u << (16 * (H & 1));
            const int w = (H - b * 8) >> 1;
            if (w == 0) pk.x = one; else if (w == 1) pk.y = one; else if (w == 2) pk.z = one; else pk.w = one;
        }
        return pk;
    };
    for (int64_t st = blockIdx.x; st < a.NT; st += gridDim.x) {
        float dh[8 * BPT];
        uint4 hq[BPT];                           // my chunks of h_{t-1}, prefetched one step ahead
        const int64_t s = st * TM + row;
#pragma unroll
        for (int u = 0; u < 8 * BPT; ++u) { const int j = blk0 * 8 + u; dh[u] = (s < a.S && j < H) ? a.dE[s * H + j] : 0.f; }
        // my chunks of the gate pre-activations: step T-1 here, step t-1 right after step t's gate epilogue
        uint4 gq[3 * BPT];
        {
            const unsigned char* g0 = reinterpret_cast<const unsigned char*>(a.ws.gi) + size_t(st * a.T + a.T - 1) * NCH * TILE_CH;
#pragma unroll
            for (int c = 0; c < 3 * BPT; ++c) gq[c] = (c < nq) ? *reinterpret_cast<const uint4*>(g0 + tile_off(TM, row, 3 * blk0 + c)) : make_uint4(0, 0, 0, 0);
        }
        for (int t = a.T - 1; t >= 0; --t) {
            if (dw_pending) { mbar_wait(&bars[2], ph2); ph2 ^= 1; dw_pending = false; }   // sHp / sDgh are free again
            // h_{t-1} operand tile
            if (t > 0) {
                if (t == a.T - 1) {
                    const unsigned char* hin = reinterpret_cast<const unsigned char*>(a.ws.hall) + size_t(st * a.T + t - 1) * HCH * TILE_CH;
#pragma unroll
                    for (int bb = 0; bb < BPT; ++bb) hq[bb] = *reinterpret_cast<const uint4*>(hin + tile_off(TM, row, blk0 + bb));
                }
            } else {
#pragma unroll
                for (int bb = 0; bb < BPT; ++bb) hq[bb] = const_chunk(blk0 + bb);
            }
#pragma unroll
            for (int bb = 0; bb < BPT; ++bb) *reinterpret_cast<uint4*>(sHp + tile_off(TM, row, blk0 + bb)) = hq[bb];
            if (blk0 == 0) for (int b = NB8; b < HCH; ++b) *reinterpret_cast<uint4*>(sHp + tile_off(TM, row, b)) = const_chunk(b);
            fence_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            if (t > 0 && warp == 0) {                         // converged issue (tc_sm100.cuh: mma_bf16_ss_w)
                tc_fence_after_sync();
                issue_row_gemm_w(tmem, 0, smem_u32(sHp), smem_u32(sWhh), NC, NC, HP / 16);
                mma_commit_w(&bars[0]);
            }
            unsigned char* gio = reinterpret_cast<unsigned char*>(a.ws.gi) + size_t(st * a.T + t) * NCH * TILE_CH;
            // (the gate math reads h_{t-1} back from sHp: carrying a second register copy across the prefetch below makes ptxas
            // rotate registers with moves placed right behind the loads, where the warp then waits out the memory latency)
            {   // next step's h_{t-2}, in flight while the MMA runs.  Unconditional, with a clamped step (t <= 1 fetches h_0 again; step 0
                // replaces hq by the constant chunk): a load under `if (t > 1)` lands in temporaries and ptxas puts the predicated
                // moves into hq right behind it -- the warp then waits out the whole memory latency there (12 % of the stall samples)
                const unsigned char* hin = reinterpret_cast<const unsigned char*>(a.ws.hall) + size_t(st * a.T + (t > 1 ? t - 2 : 0)) * HCH * TILE_CH;
#pragma unroll
                for (int bb = 0; bb < BPT; ++bb) hq[bb] = *reinterpret_cast<const uint4*>(hin + tile_off(TM, row, blk0 + bb));
            }
            if (t > 0) {
                mbar_wait(&bars[0], ph0);
                ph0 ^= 1;
                tc_fence_after_sync();
            }
#pragma unroll
            for (int bb = 0; bb < BPT; ++bb) {
                const int blk = blk0 + bb;
                const bool cblk = cmp && blk == NB8 - 1;
                float ghr[8], ghz[8], ghn[8], gir[8], giz[8], gin8[8], hp[8];
                if (t > 0) {
                    if (cblk) {
                        float a8[8], b8[8];
                        tmem_ld8x2(tmem_addr(tmem, lane_base, blk * 24), tmem_addr(tmem, lane_base, blk * 24 + 8), a8, b8);
                        expand_compact(a8, b8, ghr, ghz, ghn);
                    } else {
                        tmem_ld8x3(tmem_addr(tmem, lane_base, blk * 24), tmem_addr(tmem, lane_base, blk * 24 + 8),
                                   tmem_addr(tmem, lane_base, blk * 24 + 16), ghr, ghz, ghn);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) ghr[u] = ghz[u] = ghn[u] = 0.f;
                }
                if (cblk) {
                    float a8[8], b8[8];
                    unpack8(gq[3 * bb], a8);
                    unpack8(gq[3 * bb + 1], b8);
                    expand_compact(a8, b8, gir, giz, gin8);
                } else {
                    unpack8(gq[3 * bb], gir);
                    unpack8(gq[3 * bb + 1], giz);
                    unpack8(gq[3 * bb + 2], gin8);
                }
                unpack8(*reinterpret_cast<const uint4*>(sHp + tile_off(TM, row, blk)), hp);
                float dar[8], daz[8], dan[8], dnr[8];
#pragma unroll
                for (int p2 = 0; p2 < 4; ++p2) {                 // two hidden units per packed instruction (tc_sm100.cuh: fma2)
                    const int u = 2 * p2, j = blk * 8 + u;
                    const float2 one = splat2(1.f), neg = splat2(-1.f);
                    const float2 r = sigmoid2(add2(make_float2(gir[u], gir[u + 1]), make_float2(ghr[u], ghr[u + 1])));
                    const float2 z = sigmoid2(add2(make_float2(giz[u], giz[u + 1]), make_float2(ghz[u], ghz[u + 1])));
                    const float2 hn = add2(make_float2(ghn[u], ghn[u + 1]), *reinterpret_cast<const float2*>(sBhn + j));
                    const float2 n = tanh2(fma2(r, hn, make_float2(gin8[u], gin8[u + 1])));
                    const float2 hprev = make_float2(j < H ? hp[u] : 0.f, j + 1 < H ? hp[u + 1] : 0.f);
                    const float2 d = make_float2(dh[bb * 8 + u], dh[bb * 8 + u + 1]);
                    const float2 omz = fma2(z, neg, one);                             // 1 - z
                    const float2 dn = mul2(d, omz);
                    const float2 dz = mul2(d, fma2(n, neg, hprev));                   // d (h_prev - n)
                    const float2 da = mul2(dn, fma2(mul2(n, neg), n, one));           // dn (1 - n^2)
                    const float2 dr = mul2(mul2(da, hn), mul2(r, fma2(r, neg, one))); // da hn r (1 - r)
                    const float2 dzz = mul2(dz, mul2(z, omz));                        // dz z (1 - z)
                    const float2 dq = mul2(da, r);
                    const float2 dhz = mul2(d, z);
                    dan[u] = da.x; dan[u + 1] = da.y;
                    dar[u] = dr.x; dar[u + 1] = dr.y;
                    daz[u] = dzz.x; daz[u + 1] = dzz.y;
                    dnr[u] = dq.x; dnr[u + 1] = dq.y;
                    dh[bb * 8 + u] = dhz.x; dh[bb * 8 + u + 1] = dhz.y;
                }
                const uint4 pr = make_uint4(pack_bf16(dar[0], dar[1]), pack_bf16(dar[2], dar[3]), pack_bf16(dar[4], dar[5]), pack_bf16(dar[6], dar[7]));
                const uint4 pz = make_uint4(pack_bf16(daz[0], daz[1]), pack_bf16(daz[2], daz[3]), pack_bf16(daz[4], daz[5]), pack_bf16(daz[6], daz[7]));
                const uint4 pn = make_uint4(pack_bf16(dan[0], dan[1]), pack_bf16(dan[2], dan[3]), pack_bf16(dan[4], dan[5]), pack_bf16(dan[6], dan[7]));
                const uint4 pq = make_uint4(pack_bf16(dnr[0], dnr[1]), pack_bf16(dnr[2], dnr[3]), pack_bf16(dnr[4], dnr[5]), pack_bf16(dnr[6], dnr[7]));
                if (cblk) {                                                                // [dr4 dz4] [dn4 0000]
                    const uint4 ca = make_uint4(pr.x, pr.y, pz.x, pz.y);
                    *reinterpret_cast<uint4*>(gio + tile_off(TM, row, 3 * blk)) = ca;
                    *reinterpret_cast<uint4*>(gio + tile_off(TM, row, 3 * blk + 1)) = make_uint4(pn.x, pn.y, 0u, 0u);
                    *reinterpret_cast<uint4*>(sDgh + tile_off(TM, row, 3 * blk)) = ca;
                    *reinterpret_cast<uint4*>(sDgh + tile_off(TM, row, 3 * blk + 1)) = make_uint4(pq.x, pq.y, 0u, 0u);
                } else {
                *reinterpret_cast<uint4*>(gio + tile_off(TM, row, 3 * blk)) = pr;          // d gi, in place
                *reinterpret_cast<uint4*>(gio + tile_off(TM, row, 3 * blk + 1)) = pz;
                *reinterpret_cast<uint4*>(gio + tile_off(TM, row, 3 * blk + 2)) = pn;
                *reinterpret_cast<uint4*>(sDgh + tile_off(TM, row, 3 * blk)) = pr;         // d gh operand tile
                *reinterpret_cast<uint4*>(sDgh + tile_off(TM, row, 3 * blk + 1)) = pz;
                *reinterpret_cast<uint4*>(sDgh + tile_off(TM, row, 3 * blk + 2)) = pq;
                }
            }
            fence_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            if (warp == 0) {
                tc_fence_after_sync();
                issue_row_gemm_w(tmem, COL_DH, smem_u32(sDgh), smem_u32(sWhhT), HP, HP, NC / 16);     // dh part = dgh . W_hh
                mma_commit_w(&bars[1]);
                for (int mb = 0; mb < MB; ++mb)                                                          // dW_hh += dgh^T [h_{t-1} | 1]
                    issue_wgrad_acc_w(tmem, COL_DW + mb * HP, smem_u32(sDgh), 16 * mb, smem_u32(sHp), HP, dw_started);
                mma_commit_w(&bars[2]);
            }
            dw_started = true;
            dw_pending = true;
            {   // gate pre-activations of step t-1 of this tile, in flight from here until the next gate epilogue.  Issued behind the
                // proxy fence + barrier (the fence's MEMBAR would wait for them) and unconditional like the h prefetch (step 0
                // re-reads what it just wrote and drops it)
                const unsigned char* gn = gio - size_t(t > 0 ? NCH : 0) * TILE_CH;
#pragma unroll
                for (int c = 0; c < 3 * BPT; ++c) if (c < nq) gq[c] = *reinterpret_cast<const uint4*>(gn + tile_off(TM, row, 3 * blk0 + c));
            }
            mbar_wait(&bars[1], ph1);
            ph1 ^= 1;
            tc_fence_after_sync();
            if constexpr (BPT == 3) {                           // one TMEM round trip for the three blocks of my row
                float v0[8], v1[8], v2[8];
                tmem_ld8x3(tmem_addr(tmem, lane_base, COL_DH + blk0 * 8), tmem_addr(tmem, lane_base, COL_DH + blk0 * 8 + 8),
                           tmem_addr(tmem, lane_base, COL_DH + blk0 * 8 + 16), v0, v1, v2);
#pragma unroll
                for (int u = 0; u < 8; ++u) { dh[u] += v0[u]; dh[8 + u] += v1[u]; dh[16 + u] += v2[u]; }
            } else {
#pragma unroll
                for (int bb = 0; bb < BPT; ++bb) {
                    float v[8];
                    tmem_ld8(tmem_addr(tmem, lane_base, COL_DH + (blk0 + bb) * 8), v);
#pragma unroll
                    for (int u = 0; u < 8; ++u) dh[bb * 8 + u] += v[u];
                }
            }
            tc_fence_before_sync();
        }
    }
    if (dw_pending) { mbar_wait(&bars[2], ph2); }
    tc_fence_after_sync();
    // flush dW_hh / db_hh: lane = permuted gate row, columns = hidden index (column H = bias); the 8-column chunks are
    // dealt round-robin to the thread blocks
    if (dw_started) {
        for (int mb = 0; mb < MB; ++mb) {
            const int col = mb * 128 + row;
            int gate, j;
            const bool ok = col < NC && unperm_col(col, H, gate, j);
            for (int c8 = tblk; c8 < HCH; c8 += NTB) {
                float v[8];
                tmem_ld8(tmem_addr(tmem, lane_base, COL_DW + mb * HP + c8 * 8), v);
                if (ok) {
                    float* grow = a.gWhh + size_t(gate * H + j) * H;
                    if ((H & 3) == 0) {                           // 16-byte aligned rows: vector reductions
#pragma unroll
                        for (int u4 = 0; u4 < 2; ++u4) {
                            const int k = c8 * 8 + 4 * u4;
                            if (k < H) red_add_v4(grow + k, v[4 * u4], v[4 * u4 + 1], v[4 * u4 + 2], v[4 * u4 + 3]);
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) if (c8 * 8 + u == H) atomicAdd(a.gbhh + gate * H + j, v[u]);
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int k = c8 * 8 + u;
                            if (k < H) atomicAdd(grow + k, v[u]);
                            else if (k == H) atomicAdd(a.gbhh + gate * H + j, v[u]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TCOLS>(tmem);
}

// ---- K5: assemble the parameter gradients from Q / dWih ---------------------------------------------------------------------
struct PostArgs {
    int C, H, NC;
    const float *ln_w, *ln_b, *W1;
    const float *q, *dwih;
    FeG g;
};
__global__ void tc_post_kernel(PostArgs a) {
    const int C = a.C, H = a.H;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    // dW1[o][i] = Q[o][i] gamma[i] + db1[o] beta[i];  db1[o] = Q[o][C].  Column C+1 is the LayerNorm-fold correction of the
    // TMA kernels, Q[o][i] = D[o][i] - D[o][C+1] (fe_tc_tma.cuh); the streaming kernels leave it at zero (xhat[:, C+1] = 0).
    for (int idx = tid; idx < C * C; idx += nth) {
        const int o = idx / C, i = idx % C;
        atomicAdd(a.g.W1 + idx, (a.q[o * CP + i] - a.q[o * CP + C + 1]) * a.ln_w[i] + a.q[o * CP + C] * a.ln_b[i]);
    }
    for (int o = tid; o < C; o += nth) atomicAdd(a.g.b1 + o, a.q[o * CP + C]);
    // dgamma[i] = sum_o W1[o][i] Q[o][i];  dbeta[i] = sum_o W1[o][i] db1[o]
    // 8 threads per column, spread over the grid: the column sums are latency-bound otherwise
    for (int w0 = tid; w0 < ((C * 8 + 31) & ~31); w0 += nth) {
        const int i = w0 >> 3, part = w0 & 7;
        float dg = 0.f, db = 0.f;
        if (i < C)
            for (int o = part; o < C; o += 8) {
                const float w = a.W1[o * C + i];
                dg = fmaf(w, a.q[o * CP + i] - a.q[o * CP + C + 1], dg);
                db = fmaf(w, a.q[o * CP + C], db);
            }
#pragma unroll
        for (int s = 4; s > 0; s >>= 1) { dg += __shfl_xor_sync(0xffffffffu, dg, s); db += __shfl_xor_sync(0xffffffffu, db, s); }
        if (i < C && part == 0) { atomicAdd(a.g.ln_w + i, dg); atomicAdd(a.g.ln_b + i, db); }
    }
    // dW_ih / db_ih: un-permute the gate rows
    for (int idx = tid; idx < 3 * H * (C + 1); idx += nth) {
        const int g = idx / (C + 1), i = idx % (C + 1);
        const int col = perm_col(g / H, g % H, H);
        const float v = a.dwih[size_t(col) * CP + (i < C ? i : C)];
        if (i < C) atomicAdd(a.g.Wih + size_t(g) * C + i, v);
        else atomicAdd(a.g.bih + g, v);
    }
}

// ---- host side ------------------------------------------------------------------------------------------------------------
constexpr size_t kMaxSmem = 227 * 1024;

int num_sms() {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}

template <typename KernelT>
int launch_smem(KernelT k, int grid, size_t smem, cudaStream_t st, const ItemArgs& a) {
    if (smem > 227 * 1024) return FVAE_ERR_LIMIT;
    cudaError_t ce = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (ce != cudaSuccess) return int(ce);
    k<<<grid, NTH, smem, st>>>(a); count_launch();
    return int(cudaGetLastError());
}
template <typename KernelT>
int launch_gru(KernelT k, int threads, uint32_t tmem_cols, size_t smem, cudaStream_t st, const GruArgs& a) {
    if (smem > 227 * 1024) return FVAE_ERR_LIMIT;
    cudaError_t ce = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (ce != cudaSuccess) return int(ce);
    // resident CTAs per SM, from the hardware limits the kernel touches: threads, registers (allocated per warp in units
    // of 256), shared memory (+1 KB reserved per CTA), TMEM columns
    cudaFuncAttributes fa;
    if ((ce = cudaFuncGetAttributes(&fa, k)) != cudaSuccess) return int(ce);
    const int regs_per_warp = ((fa.numRegs * 32 + 255) / 256) * 256;
    int occ = 2048 / threads;
    const int by_regs = 65536 / (regs_per_warp * (threads / 32));
    const int by_smem = int(size_t(228 * 1024) / (smem + fa.sharedSizeBytes + 1024));      // 228 KB per SM, 1 KB reserved per CTA
    const int by_tmem = int(512u / tmem_cols);
    if (occ > by_regs) occ = by_regs;
    if (occ > by_smem) occ = by_smem;
    if (occ > by_tmem) occ = by_tmem;
    if (occ < 1) occ = 1;
    const int64_t slots = int64_t(num_sms()) * occ;
    const int grid = int(a.NT < slots ? a.NT : slots);
    k<<<grid, threads, smem, st>>>(a); count_launch();
    return int(cudaGetLastError());
}

ItemArgs make_item_args(const FeDims& d, const fvae_panel& x, const TcWs& ws) {
    ItemArgs a;
    a.x = x.data; a.seq_pitch = x.seq_pitch; a.row_pitch = x.row_pitch; a.row_index = x.row_index; a.num_rows = x.num_rows;
    a.S = d.S; a.T = d.T; a.C = d.C; a.H = d.H; a.NC = nc_of(d.H); a.HP = hp_of(d.H);
    a.NT = (int64_t(d.S) + TM - 1) / TM;
    a.items32 = (d.T >= 2 && a.NT * int64_t(d.T) * int64_t(d.T) < (int64_t(1) << 32)) ? 1 : 0;
    a.t_magic = d.T >= 2 ? uint32_t((uint64_t(1) << 32) / uint64_t(d.T)) + 1u : 0u;
    a.prefetch = 0;
    a.ws = ws;
    return a;
}

#define FVAE_DISPATCH_NB8(NB, CALL)                                   \
    switch (NB) {                                                     \
        case 1: { constexpr int kNB = 1; CALL; } break;               \
        case 2: { constexpr int kNB = 2; CALL; } break;               \
        case 3: { constexpr int kNB = 3; CALL; } break;               \
        case 4: { constexpr int kNB = 4; CALL; } break;               \
        case 5: { constexpr int kNB = 5; CALL; } break;               \
        case 6: { constexpr int kNB = 6; CALL; } break;               \
        case 7: { constexpr int kNB = 7; CALL; } break;               \
        default: { constexpr int kNB = 8; CALL; } break;              \
    }

}  // namespace

int fe_tc_supported(const FeDims& d) {
    if (d.C >= CP || d.C < 8 || d.H > kMaxH || d.H < 1) return FVAE_ERR_UNSUPPORTED;   // column C carries the constant 1
    return 0;
}

int64_t fe_tc_workspace_bytes(const FeDims& d) { return carve_tc(d, nullptr).bytes; }

// K1 alone (the operand images must already be in the workspace): used by fe_tc_forward and by bench.py's
// per-kernel roofline timing through fvae_debug_front_forward.
int fe_tc_front_only(const FeDims& d, const fvae_panel& x, void* wsp, cudaStream_t st) {
    TcWs ws = carve_tc(d, wsp);
    const int NC = nc_of(d.H);
    ItemArgs a = make_item_args(d, x, ws);
    const int nsm = num_sms();
    const int64_t nitems = a.NT * d.T;
    const int grid = int(nitems < nsm ? nitems : nsm);
    if (tma_panel_ok(x, d)) {
        XMaps xm;
        if (make_x_maps(&xm, x, d)) {
            TmaFrontArgs ta{d.T, d.C, NC, a.NT, d.S, x.row_index, int32_t(x.num_rows), getenv("FVAE_TIMELINE") ? atoi(getenv("FVAE_TIMELINE")) : 0, ws};
            const size_t fixed = A_BYTES + W1_BYTES + size_t(KCH) * NC * 16 + 2 * CP * 4 + 4 * TM * 8 + 256 + 1024;
            const int xst = (fixed + 2 * XSTAGE <= kMaxSmem) ? 2 : 1;
            const int ngi = (320 + 2 * NC <= 512) ? 2 : 1;
            const size_t smem_t = fixed + size_t(xst) * XSTAGE;
            if (smem_t <= kMaxSmem && 320 + NC <= 512) {
                auto go = [&](auto kern) -> int {
                    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_t));
                    if (ce != cudaSuccess) return int(ce);
                    kern<<<grid, TF_THREADS, smem_t, st>>>(xm, ta); count_launch();
                    return int(cudaGetLastError());
                };
                const bool idx = x.row_index != nullptr, xh = !tma_fused_backward_ok(x, d);
#define FVAE_TF(XS, NG) (idx ? (xh ? go(tc_front_tma_kernel<XS, NG, true, true>) : go(tc_front_tma_kernel<XS, NG, false, true>)) \
                             : (xh ? go(tc_front_tma_kernel<XS, NG, true, false>) : go(tc_front_tma_kernel<XS, NG, false, false>)))
                if (xst == 2 && ngi == 2) return FVAE_TF(2, 2);
                if (xst == 2) return FVAE_TF(2, 1);
                if (ngi == 2) return FVAE_TF(1, 2);
                return FVAE_TF(1, 1);
#undef FVAE_TF
            }
        }
    }
    const size_t tail = (CP + NC + 2 * NSPLIT * TM) * 4 + TM + 64;
    const size_t with_stage = W1_BYTES + size_t(KCH) * NC * 16 + 2 * A_BYTES + STAGE_BYTES + tail;
    a.prefetch = (x.dtype == FVAE_BF16 && with_stage <= kMaxSmem) ? 1 : 0;
    const size_t smem = a.prefetch ? with_stage : W1_BYTES + size_t(KCH) * NC * 16 + A_BYTES + STAGE_BYTES + tail;
    const bool idx = x.row_index != nullptr;
    if (x.dtype == FVAE_BF16) {
        if (a.prefetch) return idx ? launch_smem(tc_front_fwd_kernel<__nv_bfloat16, true, true>, grid, smem, st, a)
                                   : launch_smem(tc_front_fwd_kernel<__nv_bfloat16, false, true>, grid, smem, st, a);
        return idx ? launch_smem(tc_front_fwd_kernel<__nv_bfloat16, true, false>, grid, smem, st, a)
                   : launch_smem(tc_front_fwd_kernel<__nv_bfloat16, false, false>, grid, smem, st, a);
    }
    return idx ? launch_smem(tc_front_fwd_kernel<float, true, false>, grid, smem, st, a)
               : launch_smem(tc_front_fwd_kernel<float, false, false>, grid, smem, st, a);
}

int fe_tc_forward(const FeDims& d, const fvae_panel& x, const FeW& w, float* e, void* wsp, cudaStream_t st) {
    TcWs ws = carve_tc(d, wsp);
    const int NC = nc_of(d.H), HP = hp_of(d.H), NB = nb8_of(d.H);
    PrepArgs p{d.C, d.H, NC, HP, w.ln_w, w.ln_b, w.W1, w.b1, w.Wih, w.Whh, w.bih, w.bhh, ws};
    tc_prep_kernel<<<64, 256, 0, st>>>(p); count_launch();
    ItemArgs a = make_item_args(d, x, ws);
    const int nsm = num_sms();
    int rc;
    if ((rc = fe_tc_front_only(d, x, wsp, st)) != 0) return rc;
    GruArgs g{d.S, d.T, d.H, NC, HP, a.NT, ws, e, nullptr, nullptr, nullptr, 0};
    size_t smem = size_t(HP / 8) * NC * 16 + size_t(HP / 8) * TILE_CH + HP * 4 + 64;
    {   // gate pre-activation ring: only while four CTAs still fit one SM
        const size_t with_ring = smem + 256 + 2 * size_t(NC / 8) * TILE_CH;
        const size_t ring_cap = NC <= 64 ? 45600 : 56 * 1024;           // five CTAs per SM with a 64-column accumulator (228 KB / 5 - 1 KB)
        if (NC <= 128 && with_ring <= ring_cap) { g.gi_ring = 1; smem = with_ring; }
    }
    if (NC <= 64) { FVAE_DISPATCH_NB8(NB, rc = (gate_compact(d.H) ? launch_gru(tc_gru_fwd_kernel<kNB, 64, true>, gru_threads(kNB), 64, smem, st, g) : launch_gru(tc_gru_fwd_kernel<kNB, 64, false>, gru_threads(kNB), 64, smem, st, g))); }
    else if (NC <= 128) { FVAE_DISPATCH_NB8(NB, rc = (gate_compact(d.H) ? launch_gru(tc_gru_fwd_kernel<kNB, 128, true>, gru_threads(kNB), 128, smem, st, g) : launch_gru(tc_gru_fwd_kernel<kNB, 128, false>, gru_threads(kNB), 128, smem, st, g))); }
    else { FVAE_DISPATCH_NB8(NB, rc = (gate_compact(d.H) ? launch_gru(tc_gru_fwd_kernel<kNB, 256, true>, gru_threads(kNB), 256, smem, st, g) : launch_gru(tc_gru_fwd_kernel<kNB, 256, false>, gru_threads(kNB), 256, smem, st, g))); }
    return rc;
}

int fe_tc_backward(const FeDims& d, const fvae_panel& x, const FeW& w, const FeG& gr, const float* dE, void* wsp, cudaStream_t st) {
    TcWs ws = carve_tc(d, wsp);
    const int NC = nc_of(d.H), HP = hp_of(d.H), NB = nb8_of(d.H);
    ItemArgs a = make_item_args(d, x, ws);
    const int nsm = num_sms();
    int rc;
    {   // BPTT
        GruArgs g{d.S, d.T, d.H, NC, HP, a.NT, ws, nullptr, dE, gr.Whh, gr.bhh, 0};
        const int MB = NC > 128 ? 2 : 1;
        const size_t smem = size_t(HP / 8) * NC * 16 + size_t(NC / 8) * HP * 16 + size_t(HP / 8) * TILE_CH + size_t(16 * MB) * TILE_CH + HP * 4 + 64;
        const uint32_t cols = uint32_t(NC + MB * HP);
        if (cols <= 128) { FVAE_DISPATCH_NB8(NB, rc = (gate_compact(d.H) ? launch_gru(tc_gru_bwd_kernel<kNB, 128, true>, gru_threads(kNB), 128, smem, st, g) : launch_gru(tc_gru_bwd_kernel<kNB, 128, false>, gru_threads(kNB), 128, smem, st, g))); }
        else if (cols <= 256) { FVAE_DISPATCH_NB8(NB, rc = (gate_compact(d.H) ? launch_gru(tc_gru_bwd_kernel<kNB, 256, true>, gru_threads(kNB), 256, smem, st, g) : launch_gru(tc_gru_bwd_kernel<kNB, 256, false>, gru_threads(kNB), 256, smem, st, g))); }
        else { FVAE_DISPATCH_NB8(NB, rc = (gate_compact(d.H) ? launch_gru(tc_gru_bwd_kernel<kNB, 512, true>, gru_threads(kNB), 512, smem, st, g) : launch_gru(tc_gru_bwd_kernel<kNB, 512, false>, gru_threads(kNB), 512, smem, st, g))); }
        if (rc != 0) return rc;
    }
    // q and dwih are adjacent in the workspace (carve_tc): one memset node
    cudaError_t ce = cudaMemsetAsync(ws.q, 0, size_t(reinterpret_cast<char*>(ws.dwih) - reinterpret_cast<char*>(ws.q)) + size_t(256) * CP * 4, st);
    if (ce != cudaSuccess) return int(ce);
    const int64_t nitems = a.NT * d.T;
    const int grid = int(nitems < nsm ? nitems : nsm);
    if (tma_fused_backward_ok(x, d)) {
        // one fused kernel: raw x rows (TMA) + dGI tiles (bulk copies) read once; GEMM1 recomputed; Q and dW_ih accumulated in TMEM
        XMaps xm;
        if (!make_x_maps(&xm, x, d)) return FVAE_ERR_UNSUPPORTED;
        TmaFrontArgs ta{d.T, d.C, NC, a.NT, d.S, x.row_index, int32_t(x.num_rows), getenv("FVAE_TIMELINE") ? atoi(getenv("FVAE_TIMELINE")) : 0, ws};
        cudaError_t ce2;
        if (getenv("FVAE_BACK_SPLIT")) {
            // two roles, one launch: even CTAs du -> dpre' -> Q^T, odd CTAs GEMM1 -> u -> dW_ih^T (fe_tc_split.cuh)
            const int grid2 = (nsm & ~1) < 2 ? 2 : (nsm & ~1);
            if (x.row_index) {
                if ((ce2 = cudaFuncSetAttribute(tc_back_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TS_SMEM))) != cudaSuccess) return int(ce2);
                tc_back_split_kernel<true><<<grid2, TS_THREADS, TS_SMEM, st>>>(xm, ta); count_launch();
            } else {
                if ((ce2 = cudaFuncSetAttribute(tc_back_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TS_SMEM))) != cudaSuccess) return int(ce2);
                tc_back_split_kernel<false><<<grid2, TS_THREADS, TS_SMEM, st>>>(xm, ta); count_launch();
            }
        } else if (x.row_index) {
            if ((ce2 = cudaFuncSetAttribute(tc_back_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TB_SMEM))) != cudaSuccess) return int(ce2);
            tc_back_tma_kernel<true><<<grid, TB_THREADS, TB_SMEM, st>>>(xm, ta); count_launch();
        } else {
            if ((ce2 = cudaFuncSetAttribute(tc_back_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TB_SMEM))) != cudaSuccess) return int(ce2);
            tc_back_tma_kernel<false><<<grid, TB_THREADS, TB_SMEM, st>>>(xm, ta); count_launch();
        }
        if ((ce2 = cudaGetLastError()) != cudaSuccess) return int(ce2);
        PostArgs pf{d.C, d.H, NC, w.ln_w, w.ln_b, w.W1, ws.q, ws.dwih, gr};
        tc_post_kernel<<<64, 256, 0, st>>>(pf); count_launch();
        return int(cudaGetLastError());
    }
    {   // Q from the saved xhat tiles / mask bits (the panel is not touched)
        const size_t g_bytes = size_t(NC / 8) * TILE_CH;
        const size_t per_stage = A_BYTES + (g_bytes > A_BYTES ? g_bytes : A_BYTES);          // xhat tile + [dGI -> dpre] tile
        const size_t wt = size_t(NC / 8) * CP * 16;
        cudaError_t ce2;
#define FVAE_LAUNCH_Q(NSTGV)                                                                                          \
        do {                                                                                                          \
            const size_t smemq = wt + NSTGV * per_stage + 64;                                                         \
            if ((ce2 = cudaFuncSetAttribute(tc_q_from_tiles_kernel<NSTGV>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smemq))) != cudaSuccess) return int(ce2); \
            tc_q_from_tiles_kernel<NSTGV><<<grid, NTH, smemq, st>>>(a); count_launch();                               \
        } while (0)
        const size_t stream22 = 2 * size_t(A_BYTES) + 2 * g_bytes + 2 * size_t(A_BYTES) + wt + 128;
        const size_t stream13 = 2 * size_t(A_BYTES) + 3 * g_bytes + A_BYTES + wt + 128;
        if (stream22 <= kMaxSmem) {
            if ((ce2 = cudaFuncSetAttribute(tc_q_stream_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(stream22))) != cudaSuccess) return int(ce2);
            tc_q_stream_kernel<2, 2><<<grid, QS_THREADS, stream22, st>>>(a); count_launch();
        } else if (stream13 <= kMaxSmem) {
            if ((ce2 = cudaFuncSetAttribute(tc_q_stream_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(stream13))) != cudaSuccess) return int(ce2);
            tc_q_stream_kernel<1, 3><<<grid, QS_THREADS, stream13, st>>>(a); count_launch();
        }
        else if (wt + 3 * per_stage + 64 <= kMaxSmem) FVAE_LAUNCH_Q(3);
        else if (wt + 2 * per_stage + 64 <= kMaxSmem) FVAE_LAUNCH_Q(2);
        else if (wt + per_stage + 64 <= kMaxSmem) FVAE_LAUNCH_Q(1);
        else return FVAE_ERR_LIMIT;
#undef FVAE_LAUNCH_Q
        if ((ce2 = cudaGetLastError()) != cudaSuccess) return int(ce2);
    }
    if (NC <= 128) {   // dWih with u rebuilt from the saved xhat tiles (forward did not store u)
        const size_t g_bytes = size_t(NC / 8) * TILE_CH;
        const size_t smemw = 2 * size_t(A_BYTES) + 2 * g_bytes + A_BYTES + W1_BYTES + 128;
        if (smemw > kMaxSmem) return FVAE_ERR_LIMIT;
        cudaError_t ce2;
        if ((ce2 = cudaFuncSetAttribute(tc_wih_recompute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smemw))) != cudaSuccess) return int(ce2);
        tc_wih_recompute_kernel<<<grid, WR_THREADS, smemw, st>>>(a); count_launch();
        if ((ce2 = cudaGetLastError()) != cudaSuccess) return int(ce2);
    } else
    {   // dWih from the u tiles saved by the forward kernel (the panel is not touched)
        const size_t per_stage = size_t(NC / 8) * TILE_CH + A_BYTES;
        cudaError_t ce2;
        int nst = int((kMaxSmem - 128) / per_stage);              // as deep a ring as shared memory allows
        if (nst > 6) nst = 6;
        if (nst < 1) return FVAE_ERR_LIMIT;
        const size_t smemw = size_t(nst) * per_stage + 128;
        if ((ce2 = cudaFuncSetAttribute(tc_wih_from_u_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smemw))) != cudaSuccess) return int(ce2);
        tc_wih_from_u_kernel<<<grid, WIH_THREADS, smemw, st>>>(a, nst); count_launch();
        if ((ce2 = cudaGetLastError()) != cudaSuccess) return int(ce2);
    }
    PostArgs p{d.C, d.H, NC, w.ln_w, w.ln_b, w.W1, ws.q, ws.dwih, gr};
    tc_post_kernel<<<64, 256, 0, st>>>(p); count_launch();
    return int(cudaGetLastError());
}

}  // namespace fvae
