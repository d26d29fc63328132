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
//   warp 0      producer   : three tensor-map loads per item (2 x [128 x 64] SWIZZLE_128B + [128 x 32] SWIZZLE_64B = 40 KB)
//   warp 1      UMMA issuer: GEMM1(k+1) is queued in front of GEMM2(k); two accumulator sets for pre, one or two for GI
//   warps 2-5   row statistics: thread = row, reads its 158 features from the swizzled stage (conflict-free), mean / rstd
//   warps 6-13  u epilogue : thread = (row, 80 columns): pre -> LayerNorm fold -> LeakyReLU -> bf16 u tile (A of GEMM2)
//   warps 14-17 GI epilogue: thread = row: accumulator + bias (folded in column C) -> bf16 GI tile in HBM
#pragma once
#include <cuda.h>

constexpr int TF_THREADS = 576;
constexpr int TF_W_STAT = 2, TF_W_EPU = 6, TF_W_EPG = 14;      // first warp of each role
constexpr uint32_t XB0 = 0, XB1 = 16384, XB2 = 32768;          // x stage: [128][128B] sw128 | [128][128B] sw128 | [128][64B] sw64
constexpr uint32_t XSTAGE = 40960;

struct TmaFrontArgs {
    int T, C, NC; int64_t NT;
    int S;
    TcWs ws;
};

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
        mma_bf16_ss(tmem + dcol, ad, bd, idesc, ks > 0 ? 1u : 0u);
        bd += bstep;
    }
}

__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}

// row statistics of one staged row (thread = row): sums over exactly C = 158 features
__device__ __forceinline__ void row_stats(const unsigned char* xs, int row, int C, float& mean, float& rstd) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < KCH; ++j) {
        const uint4 p = *reinterpret_cast<const uint4*>(xs + xs_chunk_off(row, j));
        float v[8];
        unpack8(p, v);
        if (j == KCH - 1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) if (8 * j + e >= C) v[e] = 0.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) { s1 += v[e]; s2 = fmaf(v[e], v[e], s2); }
    }
    const float inv_c = 1.f / float(C);
    mean = s1 * inv_c;
    rstd = rsqrtf(fmaxf(s2 * inv_c - mean * mean, 0.f) + kLnEps);
}

// ---- K1 (TMA form) --------------------------------------------------------------------------------------------------
// XST x stages (2 when shared memory allows), NGI GI accumulator sets (2 when 320 + 2 NC <= 512 TMEM columns).
// SAVE: also write the normalised xhat tile and the LeakyReLU' sign bits the cp.async-era backward kernels stream
// (transitional / cross-check mode; the fused backward needs neither).
template <int XST, int NGI, bool SAVE>
__global__ void __launch_bounds__(TF_THREADS, 1) tc_front_tma_kernel(const __grid_constant__ CUtensorMap map128,
                                                                     const __grid_constant__ CUtensorMap map64, TmaFrontArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = a.C, NC = a.NC, NCH = NC / 8;
    unsigned char* sX = smem;                                  // [XST][XSTAGE]      (1024-byte aligned)
    unsigned char* sU = sX + XST * XSTAGE;                     // u tile [20][128][16]
    unsigned char* sW1 = sU + A_BYTES;                         // W1n image [20][160][16]
    unsigned char* sWih = sW1 + W1_BYTES;                      // W_ih image [20][NC][16]  (bias in column C)
    float* sB1 = reinterpret_cast<float*>(sWih + uint32_t(KCH) * NC * 16);   // b1f[160]
    float* sW1s = sB1 + CP;                                    // w1s[160]
    float2* sStat = reinterpret_cast<float2*>(sW1s + CP);      // [4][128] (-mean rstd, rstd)
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
    for (int i = tid; i < CP; i += TF_THREADS) { sB1[i] = a.ws.b1f[i]; sW1s[i] = a.ws.w1s[i]; }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 5); mbar_init(&pre_full[i], 1); mbar_init(&pre_empty[i], 8);
                                      mbar_init(&gi_full[i], 1); mbar_init(&gi_empty[i], 4); }
        for (int i = 0; i < 4; ++i) mbar_init(&st_full[i], 4);
        mbar_init(u_full, 8); mbar_init(u_empty, 1);
        mbar_fence_init();
        prefetch_tmap(&map128); prefetch_tmap(&map64);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    constexpr uint32_t COL_PRE = 0, COL_GI = 320;
    const int64_t nitems = a.NT * a.T, G = gridDim.x;
    const int64_t mine = nitems > int64_t(blockIdx.x) ? (nitems - 1 - blockIdx.x) / G + 1 : 0;

    if (warp == 0) {
        // ===== producer =====
        if (lane == 0) {
            for (int64_t k = 0; k < mine; ++k) {
                const int s = int(k % XST);
                if (k >= XST) mbar_wait_site(&x_empty[s], uint32_t((k / XST) - 1) & 1u, 1);
                const int64_t item = int64_t(blockIdx.x) + k * G;
                const int64_t st = item / a.T;
                const int t = int(item - st * a.T);
                unsigned char* dst = sX + s * XSTAGE;
                mbar_expect_tx(&x_full[s], XSTAGE);
                tma_load_3d(dst + XB0, &map128, 0, t, int(st * TM), &x_full[s]);
                tma_load_3d(dst + XB1, &map128, 64, t, int(st * TM), &x_full[s]);
                tma_load_3d(dst + XB2, &map64, 128, t, int(st * TM), &x_full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== UMMA issuer =====
        if (lane == 0 && mine > 0) {
            auto gemm1 = [&](int64_t k) {
                const int s = int(k % XST), b = int(k & 1);
                if (k >= 2) mbar_wait_site(&pre_empty[b], uint32_t((k >> 1) - 1) & 1u, 2);
                mbar_wait_site(&x_full[s], uint32_t(k / XST) & 1u, 3);
                tc_fence_after_sync();
                issue_gemm1_tma(tmem, COL_PRE + uint32_t(b) * CP, smem_u32(sX + s * XSTAGE), smem_u32(sW1));
                mma_commit(&pre_full[b]);
                mma_commit(&x_empty[s]);
            };
            gemm1(0);
            for (int64_t k = 0; k < mine; ++k) {
                if (k + 1 < mine) gemm1(k + 1);
                const int g = int(k % NGI);
                if (k >= NGI) mbar_wait_site(&gi_empty[g], uint32_t((k / NGI) - 1) & 1u, 4);
                mbar_wait_site(u_full, uint32_t(k) & 1u, 5);
                tc_fence_after_sync();
                issue_row_gemm(tmem, COL_GI + uint32_t(g) * NC, smem_u32(sU), smem_u32(sWih), NC, NC, KCH / 2);
                mma_commit(&gi_full[g]);
                mma_commit(u_empty);
            }
        }
    } else if (warp < TF_W_EPU) {
        // ===== row statistics =====
        const int row = (warp - TF_W_STAT) * 32 + lane;
        for (int64_t k = 0; k < mine; ++k) {
            const int s = int(k % XST), q = int(k & 3);
            mbar_wait_site(&x_full[s], uint32_t(k / XST) & 1u, 6);
            const unsigned char* xs = sX + s * XSTAGE;
            float mean, rstd;
            row_stats(xs, row, C, mean, rstd);
            sStat[q * TM + row] = make_float2(-mean * rstd, rstd);
            if (SAVE) {          // the normalised tile (column C = 1) for the streaming backward kernels
                const int64_t item = int64_t(blockIdx.x) + k * G;
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
            mbar_wait_site(&st_full[q], uint32_t(k >> 2) & 1u, 7);
            const float2 st2 = sStat[q * TM + row];
            mbar_wait_site(&pre_full[b], uint32_t(k >> 1) & 1u, 8);
            tc_fence_after_sync();
            uint4 pk[10];
            unsigned long long mlo = 0ull, mhi = 0ull;             // sign bits of my columns [0,40) and [40,80)
#pragma unroll
            for (int gq = 0; gq < 5; ++gq) {
                float v[16];
                tmem_ld16(tmem_addr(tmem, lane_base, COL_PRE + uint32_t(b) * CP + c0 + gq * 16), v);
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                    const float4 bb = *reinterpret_cast<const float4*>(sB1 + c0 + gq * 16 + 4 * e4);
                    const float4 ww = *reinterpret_cast<const float4*>(sW1s + c0 + gq * 16 + 4 * e4);
                    v[4 * e4 + 0] = fmaf(v[4 * e4 + 0], st2.y, fmaf(st2.x, ww.x, bb.x));
                    v[4 * e4 + 1] = fmaf(v[4 * e4 + 1], st2.y, fmaf(st2.x, ww.y, bb.y));
                    v[4 * e4 + 2] = fmaf(v[4 * e4 + 2], st2.y, fmaf(st2.x, ww.z, bb.z));
                    v[4 * e4 + 3] = fmaf(v[4 * e4 + 3], st2.y, fmaf(st2.x, ww.w, bb.w));
                }
                if (SAVE) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int cc = gq * 16 + e;
                        if (v[e] > 0.f) { if (cc < 40) mlo |= 1ull << cc; else mhi |= 1ull << (cc - 40); }
                    }
                }
                uint32_t w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = lrelu_pack(v[2 * e], v[2 * e + 1]);
                // the ones column (u[:, C] = 1: the GI bias rides in column C of the W_ih image), zeros beyond
                if (c0 + gq * 16 + 16 > C) {
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
            if (k > 0) mbar_wait_site(u_empty, uint32_t(k - 1) & 1u, 9);   // GEMM2(k-1) has read the u tile
#pragma unroll
            for (int ch = 0; ch < 10; ++ch) *reinterpret_cast<uint4*>(sU + tile_off(TM, row, 10 * half + ch)) = pk[ch];
            if (SAVE) {
                const int64_t item = int64_t(blockIdx.x) + k * G;
                unsigned long long* gm = a.ws.mask + size_t(item) * 4 * TM;
                gm[(2 * half) * TM + row] = mlo;
                gm[(2 * half + 1) * TM + row] = mhi;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(u_full);
        }
    } else {
        // ===== GI epilogue: thread = row =====
        const uint32_t lane_base = uint32_t(warp & 3) * 32u;
        const int row = int(lane_base) + lane;
        for (int64_t k = 0; k < mine; ++k) {
            const int g = int(k % NGI);
            const int64_t item = int64_t(blockIdx.x) + k * G;
            unsigned char* gout = reinterpret_cast<unsigned char*>(a.ws.gi) + size_t(item) * NCH * TILE_CH;
            mbar_wait_site(&gi_full[g], uint32_t(k / NGI) & 1u, 10);
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
    if (warp == 1) tmem_dealloc<512>(tmem);
}

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
// can this panel take the TMA kernels?  bf16, dense windows, every row 16-byte aligned (pitches multiples of 8 elements)
inline bool tma_panel_ok(const fvae_panel& x, const FeDims& d) {
    return x.dtype == FVAE_BF16 && x.row_index == nullptr && d.C > 128 && d.C < CP && (x.row_pitch & 7) == 0 && (x.seq_pitch & 7) == 0 &&
           (reinterpret_cast<uintptr_t>(x.data) & 15u) == 0 && x.row_pitch >= d.C && x.seq_pitch >= int64_t(d.T - 1) * x.row_pitch + d.C &&
           get_encode_tiled() != nullptr;
}
