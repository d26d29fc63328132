// Warp-specialised, TMA-fed front kernels of the tensor-core FeatureExtractor (included by fe_tc.cu inside namespace
// fvae::<anon>, after fe_tc_front.cuh).  Dense bf16 panels only (x[S][T][C] contiguous per sequence); every other panel
// form keeps the cp.async kernels of fe_tc_front.cuh.
//
// The two ideas:
//  (1) LayerNorm AFTER the GEMM.  xhat = (x - mean) rstd is linear in x per row, so
//          pre = xhat . W1g^T + b1f = rstd (x . W1g^T) - rstd mean w1s + b1f,      w1s[o] = sum_c W1g[o][c]
//      The tensor core contracts the RAW bf16 rows exactly as TMA lands them (the panel is bf16 already: no operand rounding
//      at all, where the xhat tile was a second rounding); mean / rstd are two per-row scalars applied in the epilogue that
//      reads the accumulator anyway.  No normalised tile is built, stored or re-read.  Backward uses the same identity:
//          Q = dpre^T [xhat | 1]  =  (dpre rstd)^T [x | 1/rstd | mean]   ->  Q[o][i] = D[o][i] - D[o][C+1],  db1[o] = D[o][C]
//  (2) TMA needs every box row to start on a 16-byte boundary (measured: an innermost coordinate that is not a multiple of
//      8 bf16 elements faults, scripts/probe/tma_probe.cu), and C = 158 bf16 rows are 316 bytes.  So the TMA kernels take
//      panels whose ROW PITCH is a multiple of 8 elements (x[S][T][160] with 158 features used: +1.3 % bytes; the resident
//      row table has had that pitch since round 1).  The panel is the 3-D tensor [S][T][C] with strides (seq_pitch,
//      row_pitch); the 128 rows of item (tile, t) are the box {64, 1, 128} at (0 | 64 | 128, t, tile*128) and land directly in
//      the 128-byte-swizzled K-major operand layout the UMMA reads.  The tensor's innermost extent is C, so columns C..159 of
//      the box are out of bounds and arrive as zeros whatever the padding holds.  Dense pitch-158 panels keep the cp.async path.
//
// Roles (576 threads, one CTA per SM, persistent over items; every hand-off is an mbarrier, no CTA barrier in steady state):
//   warps 0-3   row statistics: thread = row, reads its 158 features from the swizzled stage (conflict-free), mean / rstd;
//               writes (1/rstd, mean) into columns C, C+1 of the stage: rows C, C+1 of the W1n image hold (b1f, -w1s), so the
//               accumulator is pre / rstd and the whole LayerNorm fold costs the epilogue ONE multiply per element
//   warps 4-11  u epilogue : thread = (row, 80 columns): u = rstd LeakyReLU(acc) -> bf16 u tile (A of GEMM2)
//   warps 12-15 GI epilogue: thread = row: accumulator (bias folded in column C) -> bf16 GI tile in HBM
//   warp 16     producer   : three tensor-map loads per item (2 x [128 x 64] SWIZZLE_128B + [128 x 32] SWIZZLE_64B = 40 KB)
//   warp 17     UMMA issuer: GEMM1(k+1) is queued in front of GEMM2(k); two accumulator sets for pre, one or two for GI
#pragma once
#include <cuda.h>

constexpr int TF_THREADS = 576;
// Warp ids matter: a scheduler picks the HIGHEST warp id among its eligible warps, so the two single-thread roles every
// other warp waits for (UMMA issue, TMA issue) sit in the last warps; an issuer in warp 1 starves behind the epilogue warps
// of its scheduler (measured: 42 UMMAs took 3.2 k cycles to issue).
constexpr int TF_W_STAT = 0, TF_W_EPU = 4, TF_W_EPG = 12, TF_W_PROD = 16, TF_W_MMA = 17;      // first warp of each role
constexpr uint32_t XB0 = 0, XB1 = 16384, XB2 = 32768;          // x stage: [128][128B] sw128 | [128][128B] sw128 | [128][64B] sw64
constexpr uint32_t XSTAGE = 40960;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar);
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// m128 / m64: dense panel: 3-D maps, boxes {64,1,128} / {32,1,128}; resident table: 2-D maps with boxes {64,1} / {32,1} (gather4).
// t128 / t64: resident table only: 2-D maps with boxes {64,128} / {32,128} for tiles whose 128 rows are CONSECUTIVE table rows.
struct XMaps { CUtensorMap m128, m64, t128, t64; };

struct TmaFrontArgs {
    int T, C, NC; int64_t NT;
    int S;
    const int32_t* row_index;     // resident panel: [S][T] rows of the table (NULL: dense windows)
    int32_t oob_row;              // a row number past the table: gathered as zeros (sequences beyond S)
    int32_t timeline;             // diagnostics (FVAE_TIMELINE=1): CTA 0 records clock64() at every hand-off into ws.xh and prints the table
    TcWs ws;
};

// gather4: four table rows (4 x 128 B or 4 x 64 B) land as four consecutive rows of a swizzled block (measured:
// scripts/probe/gather4_probe.cu; the tensor map's box is {cols, 1})
__device__ __forceinline__ void tma_gather4(void* smem_dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}

// The x producer of both TMA kernels (whole warp 0).  Dense windows: lane 0 issues three box loads per item.  Resident
// panel: lane l gathers rows 4l..4l+3 of the tile (three gather4 per lane; the table rows of the NEXT item are fetched
// while this item's copies fly).  `release(k)` blocks until stage k % XST may be overwritten.
template <int XST, bool IDX, typename Release>
__device__ __forceinline__ void produce_x(const TmaFrontArgs& a, const XMaps* maps, unsigned char* sX,
                                          uint64_t* x_full, int64_t mine, int64_t first, int64_t G, Release release) {
    const int lane = threadIdx.x & 31;
    int4 nxt = make_int4(0, 0, 0, 0);
    auto fetch_idx = [&](int64_t k) {
        const int64_t item = first + k * G;
        const int64_t st = item / a.T;
        const int t = int(item - st * a.T);
        int r[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t sq = st * TM + 4 * lane + i;
            r[i] = sq < a.S ? a.row_index[sq * a.T + t] : a.oob_row;
        }
        return make_int4(r[0], r[1], r[2], r[3]);
    };
    if (IDX && mine > 0) nxt = fetch_idx(0);
    for (int64_t k = 0; k < mine; ++k) {
        const int s = int(k % XST);
        if (k >= XST) release(k);
        unsigned char* dst = sX + s * XSTAGE;
        if (!IDX) {
            if (lane == 0) {
                const int64_t item = first + k * G;
                const int64_t st = item / a.T;
                const int t = int(item - st * a.T);
                mbar_expect_tx(&x_full[s], XSTAGE);
                tma_load_3d(dst + XB0, &maps->m128, 0, t, int(st * TM), &x_full[s]);
                tma_load_3d(dst + XB1, &maps->m128, 64, t, int(st * TM), &x_full[s]);
                tma_load_3d(dst + XB2, &maps->m64, 128, t, int(st * TM), &x_full[s]);
            }
        } else {
            const int4 cur = nxt;
            // A tile of a dense date maps to 128 CONSECUTIVE table rows (full membership, no fill): three box loads then do what
            // 96 gather4 do otherwise -- the gather path costs the producer ~1 us per item (measured through e2e)
            const int base = __shfl_sync(0xffffffffu, cur.x, 0);
            const bool mine_ok = cur.x == base + 4 * lane && cur.y == cur.x + 1 && cur.z == cur.x + 2 && cur.w == cur.x + 3;
            const bool contiguous = __all_sync(0xffffffffu, mine_ok);
            if (lane == 0) mbar_expect_tx(&x_full[s], XSTAGE);
            __syncwarp();
            if (contiguous) {
                if (lane == 0) {
                    tma_load_2d(dst + XB0, &maps->t128, 0, base, &x_full[s]);
                    tma_load_2d(dst + XB1, &maps->t128, 64, base, &x_full[s]);
                    tma_load_2d(dst + XB2, &maps->t64, 128, base, &x_full[s]);
                }
            } else {
                tma_gather4(dst + XB0 + lane * 512, &maps->m128, 0, cur.x, cur.y, cur.z, cur.w, &x_full[s]);
                tma_gather4(dst + XB1 + lane * 512, &maps->m128, 64, cur.x, cur.y, cur.z, cur.w, &x_full[s]);
                tma_gather4(dst + XB2 + lane * 256, &maps->m64, 128, cur.x, cur.y, cur.z, cur.w, &x_full[s]);
            }
            if (k + 1 < mine) nxt = fetch_idx(k + 1);
        }
    }
}

// ---- descriptors / loads -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t make_smem_desc_sw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = make_smem_desc(saddr, lbo_bytes, sbo_bytes);
    d |= uint64_t(layout_type & 7u) << 61;          // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
    return d;
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// byte offset of 16-byte chunk j (8 features) of row r inside an x stage
__device__ __forceinline__ uint32_t xs_chunk_off(uint32_t r, uint32_t j) {
    if (j < 16) return (j >> 3) * 16384u + r * 128u + (((j & 7u) ^ (r & 7u)) << 4);
    return XB2 + r * 64u + ((((j - 16u) & 3u) ^ ((r >> 1) & 3u)) << 4);
}
// GEMM1 over an x stage: D[128 x 160] = x(raw, K-major swizzled) . W1n(K-major, chunk-major image)^T, K = 160
__device__ __forceinline__ void issue_gemm1_tma(uint32_t tmem, uint32_t dcol, uint32_t xs, uint32_t w1) {
    const uint32_t idesc = make_idesc_bf16(kTileRows, CP, false, false);
    uint64_t bd = make_smem_desc(w1, CP * 16, 128);
    const uint64_t bstep = (2 * CP * 16) >> 4;
#pragma unroll
    for (int ks = 0; ks < KCH / 2; ++ks) {
        uint64_t ad;
        if (ks < 4) ad = make_smem_desc_sw(xs + XB0 + ks * 32, 16, 1024, 2);
        else if (ks < 8) ad = make_smem_desc_sw(xs + XB1 + (ks - 4) * 32, 16, 1024, 2);
        else ad = make_smem_desc_sw(xs + XB2 + (ks - 8) * 32, 16, 512, 4);
        mma_bf16_ss_w(tmem + dcol, ad, bd, idesc, ks > 0 ? 1u : 0u);
        bd += bstep;
    }
}


// row statistics of one staged row (thread = row): sums over exactly C = 158 features
__device__ __forceinline__ void row_stats(const unsigned char* xs, int row, int C, float& mean, float& rstd) {
    // a 32-bit word holds two bf16 features = one fp32 pair (lo << 16, hi & 0xffff0000): sum and sum of squares run on packed
    // pairs (FADD2 / FFMA2), half the fma-pipe slots of the scalar form
    float2 s1 = make_float2(0.f, 0.f), s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < KCH; ++j) {
        const uint4 p = *reinterpret_cast<const uint4*>(xs + xs_chunk_off(row, j));
        uint32_t w[4] = {p.x, p.y, p.z, p.w};
        if (j == KCH - 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (8 * j + 2 * i >= C) w[i] = 0u;
                else if (8 * j + 2 * i + 1 >= C) w[i] &= 0xFFFFu;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 v = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xFFFF0000u));
            s1 = add2(s1, v);
            s2 = fma2(v, v, s2);
        }
    }
    const float inv_c = 1.f / float(C);
    mean = (s1.x + s1.y) * inv_c;
    rstd = rsqrtf(fmaxf((s2.x + s2.y) * inv_c - mean * mean, 0.f) + kLnEps);
}

// ---- K1 (TMA form) --------------------------------------------------------------------------------------------------
// XST x stages (2 when shared memory allows), NGI GI accumulator sets (2 when 320 + 2 NC <= 512 TMEM columns).
// Always saved for backward: the LeakyReLU' sign bits (40 per thread part) and the row statistics (-mean rstd, rstd).
// SAVE_XH: also write the normalised xhat tile the cp.async-era backward kernels stream (shapes the fused backward
// does not cover); IDX: rows come from the resident table through fvae_panel.row_index.
template <int XST, int NGI, bool SAVE_XH, bool IDX>
__global__ void __launch_bounds__(TF_THREADS, 1) tc_front_tma_kernel(const __grid_constant__ XMaps maps, TmaFrontArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = a.C, NC = a.NC, NCH = NC / 8;
    unsigned char* sX = smem;                                  // [XST][XSTAGE]      (1024-byte aligned)
    unsigned char* sU = sX + XST * XSTAGE;                     // u tile [20][128][16]
    unsigned char* sW1 = sU + A_BYTES;                         // W1n image [20][160][16]
    unsigned char* sWih = sW1 + W1_BYTES;                      // W_ih image [20][NC][16]  (bias in column C)
    float2* sStat = reinterpret_cast<float2*>(sWih + uint32_t(KCH) * NC * 16 + 2 * CP * 4);      // [4][128] (-mean rstd, rstd)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStat + 4 * TM);
    uint64_t* x_full = bars;            // [2]
    uint64_t* x_empty = bars + 2;       // [2]  GEMM1 commit + 4 stats warps
    uint64_t* st_full = bars + 4;       // [4]  4 stats warps
    uint64_t* pre_full = bars + 8;      // [2]
    uint64_t* pre_empty = bars + 10;    // [2]  8 epilogue warps
    uint64_t* u_full = bars + 12;       //      8 epilogue warps
    uint64_t* u_empty = bars + 13;      //      GEMM2 commit
    uint64_t* gi_full = bars + 14;      // [2]
    uint64_t* gi_empty = bars + 16;     // [2]  4 GI warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    copy_image(sW1, a.ws.w1n, W1_BYTES);
    copy_image(sWih, a.ws.wih, uint32_t(KCH) * NC * 16);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 5); mbar_init(&pre_full[i], 1); mbar_init(&pre_empty[i], 8);
                                      mbar_init(&gi_full[i], 1); mbar_init(&gi_empty[i], 4); }
        for (int i = 0; i < 4; ++i) mbar_init(&st_full[i], 4);
        mbar_init(u_full, 8); mbar_init(u_empty, 1);
        mbar_fence_init();
        prefetch_tmap(&maps.m128); prefetch_tmap(&maps.m64);
    }
    if (warp == TF_W_MMA) tmem_alloc<512>(tmem_slot);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t COL_PRE = 0, COL_GI = 320;
    const int64_t nitems = a.NT * a.T, G = gridDim.x;
    const int64_t mine = nitems > int64_t(blockIdx.x) ? (nitems - 1 - blockIdx.x) / G + 1 : 0;

    if (warp == TF_W_PROD) {
        // ===== producer =====
        if (IDX || lane == 0)
            produce_x<XST, IDX>(a, &maps, sX, x_full, mine, int64_t(blockIdx.x), G,
                                [&](int64_t k) { mbar_wait_relaxed(&x_empty[k % XST], uint32_t((k / XST) - 1) & 1u, 1); });
    } else if (warp == TF_W_MMA) {
        // ===== UMMA issuer (the whole warp runs converged; one elected lane issues: see mma_bf16_ss_w) =====
        if (mine > 0) {
            auto gemm1 = [&](int64_t k) {
                const int s = int(k % XST), b = int(k & 1);
                if (k >= 2) mbar_wait_site(&pre_empty[b], uint32_t((k >> 1) - 1) & 1u, 2);
                mbar_wait_site(&st_full[k & 3], uint32_t(k >> 2) & 1u, 3);       // landed AND columns C, C+1 = (1/rstd, mean) written
                tc_fence_after_sync();
                issue_gemm1_tma(tmem, COL_PRE + uint32_t(b) * CP, smem_u32(sX + s * XSTAGE), smem_u32(sW1));
                mma_commit_w(&pre_full[b]);
                mma_commit_w(&x_empty[s]);
            };
            gemm1(0);
            for (int64_t k = 0; k < mine; ++k) {
                if (k + 1 < mine) gemm1(k + 1);
                const int g = int(k % NGI);
                if (k >= NGI) mbar_wait_site(&gi_empty[g], uint32_t((k / NGI) - 1) & 1u, 4);
                mbar_wait_site(u_full, uint32_t(k) & 1u, 5);
                tc_fence_after_sync();
                issue_row_gemm_w(tmem, COL_GI + uint32_t(g) * NC, smem_u32(sU), smem_u32(sWih), NC, NC, KCH / 2);
                mma_commit_w(&gi_full[g]);
                mma_commit_w(u_empty);
            }
        }
    } else if (warp < TF_W_EPU) {
        // ===== row statistics =====
        const int row = (warp - TF_W_STAT) * 32 + lane;
        for (int64_t k = 0; k < mine; ++k) {
            const int s = int(k % XST), q = int(k & 3);
            mbar_wait_relaxed(&x_full[s], uint32_t(k / XST) & 1u, 6);
            const unsigned char* xs = sX + s * XSTAGE;
            float mean, rstd;
            row_stats(xs, row, C, mean, rstd);
            sStat[q * TM + row] = make_float2(-mean * rstd, rstd);
            // columns C, C+1 of the stage (the last 4 bytes of chunk 19): they meet rows (b1f, -w1s) of the W1n image
            *reinterpret_cast<uint32_t*>(const_cast<unsigned char*>(xs) + XB2 + uint32_t(row) * 64u + ((3u ^ ((uint32_t(row) >> 1) & 3u)) << 4) + 12u) =
                pack_bf16(1.f / rstd, mean);
            fence_async_smem();
            const int64_t item = int64_t(blockIdx.x) + k * G;
            a.ws.stats[size_t(item) * TM + row] = make_float2(-mean * rstd, rstd);     // saved for the fused backward
            if (SAVE_XH) {       // the normalised tile (column C = 1) for the streaming backward kernels
                unsigned char* g = reinterpret_cast<unsigned char*>(a.ws.xh) + size_t(item) * A_BYTES;
                const float shift = -mean * rstd;
#pragma unroll 4
                for (int j = 0; j < KCH; ++j) {
                    const uint4 p = *reinterpret_cast<const uint4*>(xs + xs_chunk_off(row, j));
                    float v[8];
                    unpack8(p, v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        v[e] = fmaf(v[e], rstd, shift);
                        if (8 * j + e >= C) v[e] = (8 * j + e == C) ? 1.f : 0.f;
                    }
                    *reinterpret_cast<uint4*>(g + tile_off(TM, row, j)) =
                        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                }
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(&st_full[q]); mbar_arrive(&x_empty[s]); }
        }
    } else if (warp < TF_W_EPG) {
        // ===== u epilogue: thread = (row, half) =====
        const int we = warp - TF_W_EPU;
        const int half = we >> 2;                                  // columns [80 half, 80 half + 80)
        const uint32_t lane_base = uint32_t(warp & 3) * 32u;
        const int row = int(lane_base) + lane;
        const int c0 = 80 * half;
        for (int64_t k = 0; k < mine; ++k) {
            const int b = int(k & 1), q = int(k & 3);
            mbar_wait_relaxed(&st_full[q], uint32_t(k >> 2) & 1u, 7);
            const float2 st2 = sStat[q * TM + row];
            mbar_wait_relaxed(&pre_full[b], uint32_t(k >> 1) & 1u, 8);
            tc_fence_after_sync();
            uint4 pk[10];
            uint32_t mw[3] = {0u, 0u, 0u};                         // sign bits of my 80 columns, 32 per word (one predicated OR-immediate each)
#pragma unroll
            for (int gq = 0; gq < 5; ++gq) {
                float v[16];
                tmem_ld16(tmem_addr(tmem, lane_base, COL_PRE + uint32_t(b) * CP + c0 + gq * 16), v);
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int cc = gq * 16 + e;                    // compile-time after unrolling
                    if (v[e] > 0.f) mw[cc >> 5] |= 1u << (cc & 31);
                }
                uint32_t w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {                      // rstd > 0: rstd lrelu(a) = lrelu(rstd a)
                    const float2 t = mul2(make_float2(v[2 * e], v[2 * e + 1]), splat2(st2.y));
                    w[e] = lrelu_pack(t.x, t.y);
                }
                // the ones column (u[:, C] = 1: the GI bias rides in column C of the W_ih image), zeros beyond
                if (c0 + gq * 16 + 16 > C) {                       // warp-uniform: the last 16-column group of the upper half only
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int col = c0 + gq * 16 + 2 * e;
                        if (col == C) w[e] = 0x3F80u;              // (1.0, 0.0): C even
                        else if (col + 1 == C) w[e] = (w[e] & 0xFFFFu) | 0x3F800000u;
                        else if (col > C) w[e] = 0u;
                    }
                }
                pk[2 * gq] = make_uint4(w[0], w[1], w[2], w[3]);
                pk[2 * gq + 1] = make_uint4(w[4], w[5], w[6], w[7]);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&pre_empty[b]);             // the accumulator set is free for GEMM1(k+2)
            if (k > 0) mbar_wait_relaxed(u_empty, uint32_t(k - 1) & 1u, 9);   // GEMM2(k-1) has read the u tile
#pragma unroll
            for (int ch = 0; ch < 10; ++ch) *reinterpret_cast<uint4*>(sU + tile_off(TM, row, 10 * half + ch)) = pk[ch];
            {
                const int64_t item = int64_t(blockIdx.x) + k * G;
                if (SAVE_XH && a.ws.u) {          // NC > 128: the streaming dW_ih kernel reads saved u tiles (it cannot rebuild them)
                    unsigned char* gu = reinterpret_cast<unsigned char*>(a.ws.u) + size_t(item) * A_BYTES;
#pragma unroll
                    for (int ch = 0; ch < 10; ++ch) *reinterpret_cast<uint4*>(gu + tile_off(TM, row, 10 * half + ch)) = pk[ch];
                }
                unsigned long long* gm = a.ws.mask + size_t(item) * 4 * TM;
                // saved format: 40 bits per (row, part of 40 columns)
                gm[(2 * half) * TM + row] = (unsigned long long)mw[0] | ((unsigned long long)(mw[1] & 0xFFu) << 32);
                gm[(2 * half + 1) * TM + row] = (unsigned long long)(mw[1] >> 8) | ((unsigned long long)mw[2] << 24);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(u_full);
        }
    } else if (warp < TF_W_PROD) {
        // ===== GI epilogue: thread = row =====
        const uint32_t lane_base = uint32_t(warp & 3) * 32u;
        const int row = int(lane_base) + lane;
        for (int64_t k = 0; k < mine; ++k) {
            const int g = int(k % NGI);
            const int64_t item = int64_t(blockIdx.x) + k * G;
            unsigned char* gout = reinterpret_cast<unsigned char*>(a.ws.gi) + size_t(item) * NCH * TILE_CH;
            mbar_wait_relaxed(&gi_full[g], uint32_t(k / NGI) & 1u, 10);
            tc_fence_after_sync();
#pragma unroll 1
            for (int c16 = 0; c16 < NC / 16; ++c16) {
                float v[16];
                tmem_ld16(tmem_addr(tmem, lane_base, COL_GI + uint32_t(g) * NC + c16 * 16), v);
                *reinterpret_cast<uint4*>(gout + tile_off(TM, row, 2 * c16)) =
                    make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                *reinterpret_cast<uint4*>(gout + tile_off(TM, row, 2 * c16 + 1)) =
                    make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&gi_empty[g]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == TF_W_MMA) tmem_dealloc<512>(tmem);
}

// ---- fused front backward (TMA form, NC <= 64) -------------------------------------------------------------------------
// One kernel instead of K4a (du -> dpre -> Q) + K4b (GEMM1 again -> u -> dW_ih): per item the raw x rows (TMA) and the dGI
// tile (bulk copy) are read ONCE; nothing but these two, the saved sign bits and the row statistics comes from HBM.
//   du   = dGI . W_ih                      [128 x 160]  -> dpre' = du * LeakyReLU'(pre) * rstd   (bf16 tile)
//   pre  = x . W1n^T (GEMM1 recomputed)    [128 x 160]  -> u = LeakyReLU(rstd pre + fold)        (bf16 tile)
//   Q   += dpre'^T [x | 1/rstd | mean]     (160 x 160: 128 x 160 direct + two transposed 128 x 32 blocks for rows o >= 128)
//   dWih^T += [u | 1]^T dGI                (160 x 64, transposed: two 128 x 64 blocks)
// TMEM (512 columns, all used): [0,160) du / pre ALTERNATE | [160,320) QA | [320,352) QB0 | [352,384) QB1 | [384,448) DW0 | [448,512) DW1.
// Shared memory (224 KB): x 2 x 40 KB | dGI 2 x 16 KB | ONE 40 KB tile that holds dpre'(k) until Q(k) has read it, then u(k)
// until dW(k) has | W1n image 50 KB | W_ih^T image 20 KB.
// Per item the tensor pipe runs  du(k) . [E: dpre'] . GEMM1(k), Q(k) . [E: u, under Q] . du(k+1), dW(k) ...  -- 2.6 k cycles of
// MMA per item against 2 x 0.6 k of epilogue on the critical path.
// Roles (640 threads): warps 0-15 epilogue, thread = (row, 40 columns); warp 16 writes the two LayerNorm columns (1/rstd, mean)
// into the landed x stage; warp 17 dGI producer; warp 18 x producer; warp 19 UMMA issuer.
constexpr int TB_THREADS = 640;
constexpr int TB_W_EPI = 0, TB_W_FIX = 16, TB_W_GPROD = 17, TB_W_XPROD = 18, TB_W_MMA = 19;   // issuers last: highest warp id wins the scheduler
constexpr int TB_NC = 64;
constexpr uint32_t TB_G_BYTES = (TB_NC / 8) * TILE_CH;            // 16384
constexpr uint32_t TB_OFF_G = 2 * XSTAGE, TB_OFF_UD = TB_OFF_G + 2 * TB_G_BYTES, TB_OFF_W1 = TB_OFF_UD + A_BYTES,
                   TB_OFF_WT = TB_OFF_W1 + W1_BYTES, TB_OFF_TAIL = TB_OFF_WT + (TB_NC / 8) * CP * 16;
constexpr size_t TB_SMEM = TB_OFF_TAIL + 2 * CP * 4 + 256 + 1024;

// weight-gradient MMA group over the 128 rows of an item: D[128 x N] (+)= A^T B, A / B given as (descriptor, per-16-row step)
__device__ __forceinline__ void issue_wgrad_desc(uint32_t tmem_col, uint64_t ad, uint64_t astep, uint64_t bd, uint64_t bstep, uint32_t N, bool acc) {
    const uint32_t idesc = make_idesc_bf16(kTileRows, N, true, true);
#pragma unroll
    for (int ks = 0; ks < int(kTileRows) / 16; ++ks) {
        mma_bf16_ss_w(tmem_col, ad, bd, idesc, (acc || ks > 0) ? 1u : 0u);
        ad += astep; bd += bstep;
    }
}

// diagnostics: event e of item k of CTA 0 (16 events per item)
#define TL(e) do { if (tl) tl[size_t(k) * 16 + (e)] = clock64(); } while (0)

template <bool IDX>
__global__ void __launch_bounds__(TB_THREADS, 1) tc_back_tma_kernel(const __grid_constant__ XMaps maps, TmaFrontArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = a.C, NC = a.NC;
    const uint32_t g_bytes = uint32_t(NC / 8) * TILE_CH;                   // the dGI tile of one item (NC <= 64 gate columns)
    unsigned char* sX = smem;
    unsigned char* sG = smem + TB_OFF_G;
    unsigned char* sUD = smem + TB_OFF_UD;
    unsigned char* sW1 = smem + TB_OFF_W1;
    unsigned char* sWT = smem + TB_OFF_WT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TB_OFF_TAIL);
    uint64_t* x_full = bars;            // [2] tx
    uint64_t* x_ready = bars + 2;       // [2] fix-up warp
    uint64_t* q_done = bars + 4;        // [2] commit: GEMM1(k) + Q(k) complete -> x stage and the dpre' tile are free
    uint64_t* g_full = bars + 6;        // [2] tx
    uint64_t* dw_done = bars + 8;       // [2] commit: dW(k) complete -> dGI stage and the u tile are free
    uint64_t* du_full = bars + 10;
    uint64_t* pre_full = bars + 11;
    uint64_t* dpre_full = bars + 12;    // 16 epilogue warps
    uint64_t* u_full = bars + 13;       // 16 epilogue warps
    uint64_t* fin = bars + 14;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

    copy_image(sW1, a.ws.w1n, W1_BYTES);
    copy_image(sWT, a.ws.wihT, uint32_t(NC / 8) * CP * 16);
    for (uint32_t i = tid; i < (2 * TB_G_BYTES + A_BYTES) / 16; i += TB_THREADS) reinterpret_cast<uint4*>(sG)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_ready[i], 1); mbar_init(&q_done[i], 1); mbar_init(&g_full[i], 1); mbar_init(&dw_done[i], 1); }
        mbar_init(du_full, 1); mbar_init(pre_full, 1); mbar_init(dpre_full, 16); mbar_init(u_full, 16); mbar_init(fin, 1);
        mbar_fence_init();
        prefetch_tmap(&maps.m128); prefetch_tmap(&maps.m64);
    }
    if (warp == TB_W_MMA) tmem_alloc<512>(tmem_slot);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t COL_ACC = 0, COL_QA = 160, COL_QB0 = 320, COL_QB1 = 352, COL_DW0 = 384, COL_DW1 = 448;
    const int64_t nitems = a.NT * a.T, G = gridDim.x;
    const int64_t mine = nitems > int64_t(blockIdx.x) ? (nitems - 1 - blockIdx.x) / G + 1 : 0;
    unsigned long long* tl = (a.timeline && blockIdx.x == 0 && lane == 0) ? reinterpret_cast<unsigned long long*>(a.ws.xh) : nullptr;

    if (warp == TB_W_XPROD) {
        // ===== producer: x stages (TMA) =====
        if (IDX || lane == 0)
            produce_x<2, IDX>(a, &maps, sX, x_full, mine, int64_t(blockIdx.x), G, [&](int64_t k) {
                // stage k & 1 was last read by GEMM1 / Q of item k - 2; the dGI stage of item k - 1 is requested here too
                mbar_wait_relaxed(&q_done[k & 1], uint32_t((k >> 1) - 1) & 1u, 21);
            });
    } else if (warp == TB_W_GPROD) {
        // ===== dGI producer (its own warp: it must not queue behind the x stage it does not depend on) =====
        if (lane == 0) {
            for (int64_t k = 0; k < mine; ++k) {
                const int g = int(k & 1);
                if (k >= 2) mbar_wait_relaxed(&dw_done[g], uint32_t((k >> 1) - 1) & 1u, 22);
                const int64_t item = int64_t(blockIdx.x) + k * G;
                mbar_expect_tx(&g_full[g], g_bytes);
                bulk_g2s(sG + g * TB_G_BYTES, reinterpret_cast<const unsigned char*>(a.ws.gi) + size_t(item) * g_bytes, g_bytes, &g_full[g]);
            }
        }
    } else if (warp == TB_W_MMA) {
        // ===== UMMA issuer (whole warp converged, one elected lane issues) =====
        if (mine > 0) {
            auto issue_du = [&](int64_t k) {
                mbar_wait_site(&g_full[k & 1], uint32_t(k >> 1) & 1u, 23);
                tc_fence_after_sync();
                issue_row_gemm_w(tmem, COL_ACC, smem_u32(sG + (k & 1) * TB_G_BYTES), smem_u32(sWT), CP, CP, NC / 16);
                mma_commit_w(du_full);
            };
            issue_du(0);
            const uint32_t ud = smem_u32(sUD);
            for (int64_t k = 0; k < mine; ++k) {
                const int s = int(k & 1);
                const uint32_t xs = smem_u32(sX + s * XSTAGE), gs = smem_u32(sG + s * TB_G_BYTES);
                TL(0);
                mbar_wait_site(dpre_full, uint32_t(k) & 1u, 24);
                TL(1);
                mbar_wait_site(&x_ready[s], uint32_t(k >> 1) & 1u, 25);
                TL(2);
                tc_fence_after_sync();
                issue_gemm1_tma(tmem, COL_ACC, xs, smem_u32(sW1));
                mma_commit_w(pre_full);
                const bool acc = k > 0;
                const uint64_t a_dp = make_smem_desc(ud, 128, kTileChunk);                          // dpre' tile, M block o < 128
                const uint64_t b_dp = make_smem_desc(ud + 16 * kTileChunk, 128, kTileChunk);        // dpre' tile, columns o >= 128 (N = 32)
                const uint64_t x128 = make_smem_desc_sw(xs + XB0, 16384, 1024, 2);                  // x columns [0,128)  MN-major, 2 atoms of 64
                const uint64_t x32 = make_smem_desc_sw(xs + XB2, 8192, 512, 4);                     // x columns [128,160) MN-major, 32-column atom
                issue_wgrad_desc(tmem + COL_QA, a_dp, 256 >> 4, x128, 2048 >> 4, 128, acc);         // Q[o<128][i<128]
                issue_wgrad_desc(tmem + COL_QA + 128, a_dp, 256 >> 4, x32, 1024 >> 4, 32, acc);     // Q[o<128][i>=128]
                issue_wgrad_desc(tmem + COL_QB0, x128, 2048 >> 4, b_dp, 256 >> 4, 32, acc);         // Q[o>=128][i<128]   (transposed: lane = i)
                issue_wgrad_desc(tmem + COL_QB1, x32, 1024 >> 4, b_dp, 256 >> 4, 32, acc);          // Q[o>=128][i>=128]  (lanes 0..31)
                mma_commit_w(&q_done[s]);
                TL(3);
                mbar_wait_site(u_full, uint32_t(k) & 1u, 26);
                TL(4);
                tc_fence_after_sync();
                if (k + 1 < mine) issue_du(k + 1);
                TL(5);
                const uint64_t a_u0 = make_smem_desc(ud, 128, kTileChunk), a_u1 = make_smem_desc(ud + 16 * kTileChunk, 128, kTileChunk);
                const uint64_t b_g = make_smem_desc(gs, 128, kTileChunk);
                issue_wgrad_desc(tmem + COL_DW0, a_u0, 256 >> 4, b_g, 256 >> 4, uint32_t(NC), acc);  // dWih^T[c<128][g]
                issue_wgrad_desc(tmem + COL_DW1, a_u1, 256 >> 4, b_g, 256 >> 4, uint32_t(NC), acc);  // dWih^T[c>=128][g]  (lanes 0..31; column C = bias)
                mma_commit_w(&dw_done[s]);
                TL(6);
            }
            mma_commit_w(fin);
        }
    } else if (warp == TB_W_FIX) {
        // ===== LayerNorm columns of the landed x stage: x[:, C] = 1/rstd (-> db1), x[:, C+1] = mean (-> the fold correction) =====
        for (int64_t k = 0; k < mine; ++k) {
            const int s = int(k & 1);
            const int64_t item = int64_t(blockIdx.x) + k * G;
            float2 st4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) st4[i] = a.ws.stats[size_t(item) * TM + 4 * lane + i];
            mbar_wait_relaxed(&x_full[s], uint32_t(k >> 1) & 1u, 27);
            unsigned char* xs = sX + s * XSTAGE;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t r = 4 * lane + i;
                const float inv = 1.f / st4[i].y;
                // columns C, C+1 = 158, 159: the last 4 bytes of chunk 19 (block 2, chunk 3)
                *reinterpret_cast<uint32_t*>(xs + XB2 + r * 64u + ((3u ^ ((r >> 1) & 3u)) << 4) + 12u) = pack_bf16(inv, -st4[i].x * inv);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&x_ready[s]);
        }
    } else {
        // ===== epilogue: thread = (row, part): 40 columns =====
        const int part = (warp - TB_W_EPI) >> 2;
        const uint32_t lane_base = uint32_t(warp & 3) * 32u;
        const int row = int(lane_base) + lane;
        const int c0 = HALF_COLS * part;
        auto ld40 = [&](float (&v)[40]) {
            uint32_t r0[16], r1[16];
            float r2[8];
            tmem_ld16_nowait(tmem_addr(tmem, lane_base, COL_ACC + c0), r0);
            tmem_ld16_nowait(tmem_addr(tmem, lane_base, COL_ACC + c0 + 16), r1);
            tmem_ld8(tmem_addr(tmem, lane_base, COL_ACC + c0 + 32), r2);          // waits for all three
#pragma unroll
            for (int e = 0; e < 16; ++e) { v[e] = __uint_as_float(r0[e]); v[16 + e] = __uint_as_float(r1[e]); }
#pragma unroll
            for (int e = 0; e < 8; ++e) v[32 + e] = r2[e];
        };
        if (warp != TB_W_EPI) tl = nullptr;
        unsigned long long mbits = 0ull;
        float2 st2 = make_float2(0.f, 1.f);
        if (mine > 0) {
            mbits = a.ws.mask[size_t(blockIdx.x) * 4 * TM + part * TM + row];
            st2 = a.ws.stats[size_t(blockIdx.x) * TM + row];
        }
        for (int64_t k = 0; k < mine; ++k) {
            unsigned long long nm = 0ull;
            float2 nst = make_float2(0.f, 1.f);
            if (k + 1 < mine) {
                const int64_t nitem = int64_t(blockIdx.x) + (k + 1) * G;
                nm = a.ws.mask[size_t(nitem) * 4 * TM + part * TM + row];
                nst = a.ws.stats[size_t(nitem) * TM + row];
            }
            uint4 pk[HALF_CH];
            {   // ---- dpre' = du * LeakyReLU'(pre) * rstd
                TL(7);
                mbar_wait_relaxed(du_full, uint32_t(k) & 1u, 28);
                TL(8);
                tc_fence_after_sync();
                float v[40];
                ld40(v);
                const uint32_t mlo = uint32_t(mbits), mhi = uint32_t(mbits >> 32);
                const float s_pos = st2.y, s_neg = kLeakySlope * st2.y;
                uint32_t wv[20];
#pragma unroll
                for (int e = 0; e < 20; ++e) {                     // per pair: two selects of the slope, ONE packed multiply, one pack
                    const float2 sc = make_float2((((2 * e < 32 ? mlo >> (2 * e) : mhi >> (2 * e - 32)) & 1u) ? s_pos : s_neg),
                                                  (((2 * e + 1 < 32 ? mlo >> (2 * e + 1) : mhi >> (2 * e + 1 - 32)) & 1u) ? s_pos : s_neg));
                    const float2 t = mul2(make_float2(v[2 * e], v[2 * e + 1]), sc);
                    wv[e] = pack_bf16(t.x, t.y);
                }
#pragma unroll
                for (int ch = 0; ch < HALF_CH; ++ch) pk[ch] = make_uint4(wv[4 * ch], wv[4 * ch + 1], wv[4 * ch + 2], wv[4 * ch + 3]);
                tc_fence_before_sync();
                TL(9);
                if (k > 0) mbar_wait_relaxed(&dw_done[(k - 1) & 1], uint32_t((k - 1) >> 1) & 1u, 29);      // dW(k-1) has read the u tile
                TL(10);
#pragma unroll
                for (int ch = 0; ch < HALF_CH; ++ch) *reinterpret_cast<uint4*>(sUD + tile_off(TM, row, HALF_CH * part + ch)) = pk[ch];
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(dpre_full);
                TL(11);
            }
            {   // ---- u = LeakyReLU(rstd pre + fold)
                mbar_wait_relaxed(pre_full, uint32_t(k) & 1u, 30);
                TL(12);
                tc_fence_after_sync();
                float v[40];
                ld40(v);
                const int one_ch = (C >= c0 && C < c0 + HALF_COLS) ? (C - c0) >> 3 : -1;               // warp-uniform: my chunk that holds column C
#pragma unroll
                for (int ch = 0; ch < HALF_CH; ++ch) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {                                           // rstd lrelu(acc)
                        const float2 t = mul2(make_float2(v[8 * ch + 2 * e], v[8 * ch + 2 * e + 1]), splat2(st2.y));
                        w[e] = lrelu_pack(t.x, t.y);
                    }
                    if (ch == one_ch) {                                      // u[:, C] = 1 (the bias row of dW_ih), zeros beyond
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int col = c0 + 8 * ch + 2 * e;
                            if (col == C) w[e] = 0x3F80u;
                            else if (col + 1 == C) w[e] = (w[e] & 0xFFFFu) | 0x3F800000u;
                            else if (col > C) w[e] = 0u;
                        }
                    }
                    pk[ch] = make_uint4(w[0], w[1], w[2], w[3]);
                }
                tc_fence_before_sync();
                TL(13);
                mbar_wait_relaxed(&q_done[k & 1], uint32_t(k >> 1) & 1u, 31);                              // Q(k) has read the dpre' tile
                TL(14);
#pragma unroll
                for (int ch = 0; ch < HALF_CH; ++ch) *reinterpret_cast<uint4*>(sUD + tile_off(TM, row, HALF_CH * part + ch)) = pk[ch];
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(u_full);
                TL(15);
            }
            mbits = nm; st2 = nst;
        }
        if (mine > 0) {
            // ---- flush the accumulators (raw sums: tc_post subtracts the fold column C+1)
            mbar_wait_relaxed(fin, 0, 32);
            tc_fence_after_sync();
            // QA: lane = o < 128, my 40 columns i
#pragma unroll 1
            for (int ch = 0; ch < HALF_CH; ++ch) {
                float d[8];
                tmem_ld8(tmem_addr(tmem, lane_base, COL_QA + c0 + ch * 8), d);
                red_add_v4(a.ws.q + size_t(row) * CP + c0 + ch * 8, d[0], d[1], d[2], d[3]);
                red_add_v4(a.ws.q + size_t(row) * CP + c0 + ch * 8 + 4, d[4], d[5], d[6], d[7]);
            }
            // QB0 / QB1: lane = i (QB1: i = 128 + lane, lanes 0..31), my 8 of the 32 columns o - 128
            for (int blk = 0; blk < 2; ++blk) {
                float d[8];
                tmem_ld8(tmem_addr(tmem, lane_base, (blk == 0 ? COL_QB0 : COL_QB1) + 8 * part), d);
                const int i = blk * 128 + row;
                if (i < CP) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int o = 128 + 8 * part + e;
                        if (o < C) atomicAdd(a.ws.q + size_t(o) * CP + i, d[e]);
                    }
                }
            }
            // DW0 / DW1: lane = c (DW1: c = 128 + lane), my 16 of the 64 gate columns g -> dwih[g][c]
            for (int blk = 0; blk < 2; ++blk) {
                float d[16];
                tmem_ld16(tmem_addr(tmem, lane_base, (blk == 0 ? COL_DW0 : COL_DW1) + 16 * part), d);
                const int c = blk * 128 + row;
                if (c < CP) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) if (16 * part + e < NC) atomicAdd(a.ws.dwih + size_t(16 * part + e) * CP + c, d[e]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == TB_W_MMA) tmem_dealloc<512>(tmem);
    if (a.timeline && blockIdx.x == 0 && tid == 0) {
        __threadfence();
        const unsigned long long* t = reinterpret_cast<const unsigned long long*>(a.ws.xh);
        const int64_t n = mine < 24 ? mine : 24;
        printf("fvae timeline (CTA 0, cycles relative to item start = MMA waits dpre): MMA: dpre_full x_ready q_issued u_full du_next dw_issued | EPI: du_wait du_full dpre_done dwdone_seen dpre_arrive pre_full u_done qdone_seen u_arrive\n");
        for (int64_t k = 0; k < n; ++k) {
            const unsigned long long t0 = t[k * 16];
            printf("item %2d:", int(k));
            for (int e = 1; e < 16; ++e) printf(" %6lld", (long long)(t[k * 16 + e] - t0));
            printf("  | next item starts +%lld\n", (long long)(t[(k + 1) * 16] - t0));
        }
    }
}
#undef TL

// ---- host: tensor maps ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}
// the dense bf16 panel as the 3-D tensor [S][T][C] (strides seq_pitch, row_pitch); box = [128 sequences][1 time step][box_cols]
inline bool make_panel_map(CUtensorMap* m, const fvae_panel& x, const FeDims& d, uint32_t box_cols, CUtensorMapSwizzle sw) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[3] = {cuuint64_t(d.C), cuuint64_t(d.T), cuuint64_t(d.S)};
    const cuuint64_t strides[2] = {cuuint64_t(x.row_pitch) * 2, cuuint64_t(x.seq_pitch) * 2};
    const cuuint32_t box[3] = {box_cols, 1, TM};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(x.data), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// the resident row table as [rows][C] (pitch row_pitch); box = {box_cols, 1}: what tile::gather4 takes
inline bool make_table_map(CUtensorMap* m, const fvae_panel& x, const FeDims& d, uint32_t box_cols, CUtensorMapSwizzle sw, uint32_t box_rows = 1) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[2] = {cuuint64_t(d.C), cuuint64_t(x.num_rows)};
    const cuuint64_t strides[1] = {cuuint64_t(x.row_pitch) * 2};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x.data), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// can this panel take the TMA kernels?  bf16, dense windows, every row 16-byte aligned (pitches multiples of 8 elements)
inline bool tma_panel_ok(const fvae_panel& x, const FeDims& d) {
    if (x.dtype != FVAE_BF16 || d.C <= 128 || d.C + 2 > CP || (x.row_pitch & 7) != 0 || x.row_pitch < d.C ||
        (reinterpret_cast<uintptr_t>(x.data) & 15u) != 0 || get_encode_tiled() == nullptr || getenv("FVAE_FRONT_CPASYNC"))
        return false;
    if (x.row_index) return x.num_rows > 0 && x.num_rows < (int64_t(1) << 31) - 1;
    return (x.seq_pitch & 7) == 0 && x.seq_pitch >= int64_t(d.T - 1) * x.row_pitch + d.C;
}
inline bool make_x_maps(XMaps* m, const fvae_panel& x, const FeDims& d) {
    if (x.row_index)
        return make_table_map(&m->m128, x, d, 64, CU_TENSOR_MAP_SWIZZLE_128B) && make_table_map(&m->m64, x, d, 32, CU_TENSOR_MAP_SWIZZLE_64B) &&
               make_table_map(&m->t128, x, d, 64, CU_TENSOR_MAP_SWIZZLE_128B, TM) && make_table_map(&m->t64, x, d, 32, CU_TENSOR_MAP_SWIZZLE_64B, TM);
    const bool ok = make_panel_map(&m->m128, x, d, 64, CU_TENSOR_MAP_SWIZZLE_128B) && make_panel_map(&m->m64, x, d, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    m->t128 = m->m128; m->t64 = m->m64;
    return ok;
}
// the fused backward covers NC <= 64 (H <= 20 with the compact gate layout); the forward kernel then saves no xhat tiles
inline bool tma_fused_backward_ok(const fvae_panel& x, const FeDims& d) { return tma_panel_ok(x, d) && nc_of(d.H) <= TB_NC && !getenv("FVAE_BACK_STREAM"); }

