// Front backward as TWO ROLES in one launch (included by fe_tc.cu after fe_tc_tma.cuh).  NC <= 64.
//
// The fused kernel of fe_tc_tma.cuh (tc_back_tma_kernel) keeps every accumulator of the front backward in one SM's 512 TMEM
// columns, which forces du and pre to ALTERNATE in one 160-column region: its per-item chain
//     du -> [epilogue dpre'] -> GEMM1 -> [epilogue u] -> du ...
// is serial (measured: 6.2 k cycles per item against 2.6 k of tensor work), and the 160 x 160 / 160 x 64 weight gradients have
// to be cut into 128 x 32 pieces whose UMMAs cost more to ISSUE (~65 cycles in situ) than to execute (16).
// Here even CTAs run role A, odd CTAs role B, over the same item sequence (item = (cta >> 1) + k * grid / 2):
//   role A  du(k) = dGI . W_ih -> [epilogue: dpre' = du * LeakyReLU'(pre) * rstd] -> Q^T += [x | 1/rstd | mean]^T dpre'
//           TMEM: du 160 | Q^T rows i < 128: 160 | Q^T rows i >= 128: 160  = 480.  Q^T is N = 160 wide: 16 UMMAs of 80 cycles per
//           item instead of 32 small ones; du(k+1) runs in front of Q(k), the epilogue of item k+1 under Q(k).
//   role B  pre(k) = x . W1n^T (recomputed) -> [epilogue: u = rstd LeakyReLU(acc)] -> dW_ih^T += [u | 1]^T dGI
//           TMEM: pre x 2 (320) | dW_ih^T 2 x 64 = 448.  GEMM1(k+1) runs in front of dW(k), the epilogue of k+1 under dW(k).
// Both roles read the same x rows (TMA) and dGI tiles (bulk copy) within a few items of each other: the second reader hits L2
// (126 MB), so DRAM still sees each byte once.  Epilogue: 16 warps, thread = (row, 40 columns); issuer / producer warps last.
//
// MEASURED (cfg2, 148 CTAs): 0.30 ms on one box, 0.36 ms on two others (the fused kernel: 0.32 ms on all three) -- both roles
// turn out to be bound by the shared-memory port, not by the issue chain: every 128 x 160 x 16 UMMA reads 9 KB of operands
// (115 B/cycle against 128), the tile stores and TMA fills compete with it, ~2.2 k cycles per item per role is the floor, and
// the L2 sharing between the roles depends on how far they drift apart.  The fused kernel stays the default
// (FVAE_BACK_SPLIT=1 selects this one); the kernel is kept as the measured alternative.
#pragma once

constexpr int TS_THREADS = 640;
constexpr int TS_W_FIX = 16, TS_W_GPROD = 17, TS_W_XPROD = 18, TS_W_MMA = 19;
constexpr uint32_t TS_G_BYTES = 16384;                       // dGI stage (NC <= 64)
constexpr uint32_t TS_OFF_G = 2 * XSTAGE;                    // 81920
constexpr uint32_t TS_OFF_T = TS_OFF_G + 2 * TS_G_BYTES;     // 114688: role A: dpre' tiles x 2 | role B: u tile, then the W1n image
constexpr uint32_t TS_OFF_W_A = TS_OFF_T + 2 * A_BYTES;      // role A: W_ih^T image (20 KB)
constexpr uint32_t TS_OFF_W_B = TS_OFF_T + A_BYTES;          // role B: W1n image (50 KB)
constexpr uint32_t TS_OFF_TAIL = TS_OFF_T + 2 * A_BYTES + 20480;       // max of both layouts: 217088
constexpr size_t TS_SMEM = TS_OFF_TAIL + 256 + 1024;

template <bool IDX>
__global__ void __launch_bounds__(TS_THREADS, 1) tc_back_split_kernel(const __grid_constant__ XMaps maps, TmaFrontArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int C = a.C, NC = a.NC;
    const bool roleA = (blockIdx.x & 1) == 0;
    const uint32_t g_bytes = uint32_t(NC / 8) * TILE_CH;
    unsigned char* sX = smem;
    unsigned char* sG = smem + TS_OFF_G;
    unsigned char* sT = smem + TS_OFF_T;                       // role A: dpre'[2]; role B: u
    unsigned char* sW = smem + (roleA ? TS_OFF_W_A : TS_OFF_W_B);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TS_OFF_TAIL);
    uint64_t* x_full = bars;            // [2] tx
    uint64_t* x_ready = bars + 2;       // [2] fix-up warp
    uint64_t* x_empty = bars + 4;       // [2] commit: the UMMAs that read the x stage are done (A: Q(k); B: GEMM1(k))
    uint64_t* g_full = bars + 6;        // [2] tx
    uint64_t* g_empty = bars + 8;       // [2] commit: A: du(k) done; B: dW(k) done
    uint64_t* acc_full = bars + 10;     // [2] commit: A: du(k) (only [0] used); B: pre(k & 1)
    uint64_t* acc_empty = bars + 12;    // [2] 16 epilogue warps: B only (pre set free for GEMM1(k+2))
    uint64_t* t_full = bars + 14;       // [2] 16 epilogue warps: the tile (A: dpre'(k & 1); B: u, only [0]) is written, the accumulator is read
    uint64_t* t_empty = bars + 16;      // [2] commit: A: Q(k) done with dpre'(k & 1); B: dW(k) done with u (only [0])
    uint64_t* fin = bars + 18;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

    if (roleA) copy_image(sW, a.ws.wihT, uint32_t(NC / 8) * CP * 16);
    else copy_image(sW, a.ws.w1n, W1_BYTES);
    for (uint32_t i = tid; i < (2 * TS_G_BYTES + 2 * A_BYTES) / 16; i += TS_THREADS) {
        if (!roleA && TS_OFF_G + i * 16 >= TS_OFF_W_B) break;                 // role B: the image starts after the single u tile
        reinterpret_cast<uint4*>(sG)[i] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&x_full[i], 1); mbar_init(&x_ready[i], 1); mbar_init(&x_empty[i], 1); mbar_init(&g_full[i], 1); mbar_init(&g_empty[i], 1);
            mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 16); mbar_init(&t_full[i], 16); mbar_init(&t_empty[i], 1);
        }
        mbar_init(fin, 1);
        mbar_fence_init();
        prefetch_tmap(&maps.m128); prefetch_tmap(&maps.m64);
    }
    if (warp == TS_W_MMA) tmem_alloc<512>(tmem_slot);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const int64_t nitems = a.NT * a.T;
    const int64_t G = gridDim.x >> 1, first = blockIdx.x >> 1;          // the host launches an even grid
    const int64_t mine = nitems > first ? (nitems - 1 - first) / G + 1 : 0;
    // diagnostics (FVAE_TIMELINE=1): CTAs 0 (role A) and 1 (role B) record clock64() at their hand-offs into ws.xh (unused by this path)
    unsigned long long* tl = (a.timeline && blockIdx.x < 2 && lane == 0 && (warp == TS_W_MMA || warp == 0))
                                 ? reinterpret_cast<unsigned long long*>(a.ws.xh) + size_t(blockIdx.x) * 4096 : nullptr;
#define TLS(e) do { if (tl) tl[size_t(k) * 16 + (e)] = clock64(); } while (0)

    if (warp == TS_W_XPROD) {
        if (IDX || lane == 0)
            produce_x<2, IDX>(a, &maps, sX, x_full, mine, first, G,
                              [&](int64_t k) { mbar_wait_relaxed(&x_empty[k & 1], uint32_t((k >> 1) - 1) & 1u, 41); });
    } else if (warp == TS_W_GPROD) {
        if (lane == 0) {
            for (int64_t k = 0; k < mine; ++k) {
                const int g = int(k & 1);
                if (k >= 2) mbar_wait_relaxed(&g_empty[g], uint32_t((k >> 1) - 1) & 1u, 42);
                const int64_t item = first + k * G;
                mbar_expect_tx(&g_full[g], g_bytes);
                bulk_g2s(sG + g * TS_G_BYTES, reinterpret_cast<const unsigned char*>(a.ws.gi) + size_t(item) * g_bytes, g_bytes, &g_full[g]);
            }
        }
    } else if (warp == TS_W_FIX) {
        // columns C, C+1 of the landed x stage = (1/rstd, mean): role A needs them as the db1 / fold-correction columns of
        // Q^T's A operand, role B as the LayerNorm fold of GEMM1 (rows C, C+1 of the W1n image are (b1f, -w1s))
        for (int64_t k = 0; k < mine; ++k) {
            const int s = int(k & 1);
            const int64_t item = first + k * G;
            float2 st4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) st4[i] = a.ws.stats[size_t(item) * TM + 4 * lane + i];
            mbar_wait_relaxed(&x_full[s], uint32_t(k >> 1) & 1u, 43);
            unsigned char* xs = sX + s * XSTAGE;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t r = 4 * lane + i;
                const float inv = 1.f / st4[i].y;
                *reinterpret_cast<uint32_t*>(xs + XB2 + r * 64u + ((3u ^ ((r >> 1) & 3u)) << 4) + 12u) = pack_bf16(inv, -st4[i].x * inv);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&x_ready[s]);
        }
    } else if (warp == TS_W_MMA) {
        if (mine > 0 && roleA) {
            // ===== role A issuer: du(k+1) in front of Q(k) =====
            constexpr uint32_t COL_DU = 0, COL_Q1 = 160, COL_Q2 = 320;
            auto issue_du = [&](int64_t k) {
                mbar_wait_site(&g_full[k & 1], uint32_t(k >> 1) & 1u, 44);
                tc_fence_after_sync();
                issue_row_gemm_w(tmem, COL_DU, smem_u32(sG + (k & 1) * TS_G_BYTES), smem_u32(sW), CP, CP, NC / 16);
                mma_commit_w(&acc_full[0]);
                mma_commit_w(&g_empty[k & 1]);
            };
            issue_du(0);
            for (int64_t k = 0; k < mine; ++k) {
                const int s = int(k & 1);
                const uint32_t xs = smem_u32(sX + s * XSTAGE), dp = smem_u32(sT + s * A_BYTES);
                TLS(0);
                mbar_wait_site(&t_full[s], uint32_t(k >> 1) & 1u, 45);           // dpre'(k) written, du(k) read out of TMEM
                TLS(1);
                tc_fence_after_sync();
                if (k + 1 < mine) issue_du(k + 1);
                TLS(2);
                mbar_wait_site(&x_ready[s], uint32_t(k >> 1) & 1u, 46);
                TLS(3);
                tc_fence_after_sync();
                const uint64_t b_dp = make_smem_desc(dp, 128, kTileChunk);                          // dpre' tile, MN-major, N = 160
                const uint64_t x128 = make_smem_desc_sw(xs + XB0, 16384, 1024, 2);                  // x columns [0,128): M = 128
                const uint64_t x32 = make_smem_desc_sw(xs + XB2, 8192, 512, 4);                     // x columns [128,160): lanes 0..31 (M = 128 reads past)
                issue_wgrad_desc(tmem + COL_Q1, x128, 2048 >> 4, b_dp, 256 >> 4, CP, k > 0);
                issue_wgrad_desc(tmem + COL_Q2, x32, 1024 >> 4, b_dp, 256 >> 4, CP, k > 0);
                mma_commit_w(&x_empty[s]);
                mma_commit_w(&t_empty[s]);
                TLS(4);
            }
            mma_commit_w(fin);
        } else if (mine > 0) {
            // ===== role B issuer: GEMM1(k+1) in front of dW(k) =====
            constexpr uint32_t COL_PRE = 0, COL_DW0 = 320, COL_DW1 = 384;
            auto gemm1 = [&](int64_t k) {
                const int s = int(k & 1);
                if (k >= 2) mbar_wait_site(&acc_empty[s], uint32_t((k >> 1) - 1) & 1u, 47);
                mbar_wait_site(&x_ready[s], uint32_t(k >> 1) & 1u, 48);
                tc_fence_after_sync();
                issue_gemm1_tma(tmem, COL_PRE + uint32_t(s) * CP, smem_u32(sX + s * XSTAGE), smem_u32(sW));
                mma_commit_w(&acc_full[s]);
                mma_commit_w(&x_empty[s]);
            };
            gemm1(0);
            const uint32_t ud = smem_u32(sT);
            for (int64_t k = 0; k < mine; ++k) {
                TLS(0);
                if (k + 1 < mine) gemm1(k + 1);
                TLS(1);
                const int g = int(k & 1);
                mbar_wait_site(&t_full[0], uint32_t(k) & 1u, 49);                 // u(k) written
                TLS(2);
                mbar_wait_site(&g_full[g], uint32_t(k >> 1) & 1u, 50);
                TLS(3);
                tc_fence_after_sync();
                const uint64_t a_u0 = make_smem_desc(ud, 128, kTileChunk), a_u1 = make_smem_desc(ud + 16 * kTileChunk, 128, kTileChunk);
                const uint64_t b_g = make_smem_desc(smem_u32(sG + g * TS_G_BYTES), 128, kTileChunk);
                issue_wgrad_desc(tmem + COL_DW0, a_u0, 256 >> 4, b_g, 256 >> 4, uint32_t(NC), k > 0);   // dWih^T[c<128][g]
                issue_wgrad_desc(tmem + COL_DW1, a_u1, 256 >> 4, b_g, 256 >> 4, uint32_t(NC), k > 0);   // dWih^T[c>=128][g] (lanes 0..31; column C = bias)
                mma_commit_w(&t_empty[0]);
                mma_commit_w(&g_empty[g]);
                TLS(4);
            }
            mma_commit_w(fin);
        }
    } else {
        // ===== epilogue: thread = (row, part): 40 columns =====
        const int part = warp >> 2;
        const uint32_t lane_base = uint32_t(warp & 3) * 32u;
        const int row = int(lane_base) + lane;
        const int c0 = HALF_COLS * part;
        auto ld40 = [&](uint32_t col, float (&v)[40]) {
            uint32_t r0[16], r1[16];
            float r2[8];
            tmem_ld16_nowait(tmem_addr(tmem, lane_base, col + c0), r0);
            tmem_ld16_nowait(tmem_addr(tmem, lane_base, col + c0 + 16), r1);
            tmem_ld8(tmem_addr(tmem, lane_base, col + c0 + 32), r2);
#pragma unroll
            for (int e = 0; e < 16; ++e) { v[e] = __uint_as_float(r0[e]); v[16 + e] = __uint_as_float(r1[e]); }
#pragma unroll
            for (int e = 0; e < 8; ++e) v[32 + e] = r2[e];
        };
        if (roleA) {
            unsigned long long mbits = 0ull;
            float rstd = 1.f;
            if (mine > 0) {
                mbits = a.ws.mask[size_t(first) * 4 * TM + part * TM + row];
                rstd = a.ws.stats[size_t(first) * TM + row].y;
            }
            for (int64_t k = 0; k < mine; ++k) {
                unsigned long long nm = 0ull;
                float nr = 1.f;
                if (k + 1 < mine) {
                    const int64_t nitem = first + (k + 1) * G;
                    nm = a.ws.mask[size_t(nitem) * 4 * TM + part * TM + row];
                    nr = a.ws.stats[size_t(nitem) * TM + row].y;
                }
                TLS(8);
                mbar_wait_relaxed(&acc_full[0], uint32_t(k) & 1u, 51);
                TLS(9);
                tc_fence_after_sync();
                float v[40];
                ld40(0, v);
                const uint32_t mlo = uint32_t(mbits), mhi = uint32_t(mbits >> 32);
                const float s_pos = rstd, s_neg = kLeakySlope * rstd;
#pragma unroll
                for (int e = 0; e < 40; ++e) v[e] *= ((e < 32 ? mlo >> e : mhi >> (e - 32)) & 1u) ? s_pos : s_neg;
                uint4 pk[HALF_CH];
#pragma unroll
                for (int ch = 0; ch < HALF_CH; ++ch)
                    pk[ch] = make_uint4(pack_bf16(v[8 * ch], v[8 * ch + 1]), pack_bf16(v[8 * ch + 2], v[8 * ch + 3]),
                                        pack_bf16(v[8 * ch + 4], v[8 * ch + 5]), pack_bf16(v[8 * ch + 6], v[8 * ch + 7]));
                tc_fence_before_sync();
                const int s = int(k & 1);
                TLS(10);
                if (k >= 2) mbar_wait_relaxed(&t_empty[s], uint32_t((k >> 1) - 1) & 1u, 52);          // Q(k-2) has read this dpre' tile
                TLS(11);
                unsigned char* dst = sT + s * A_BYTES;
#pragma unroll
                for (int ch = 0; ch < HALF_CH; ++ch) *reinterpret_cast<uint4*>(dst + tile_off(TM, row, HALF_CH * part + ch)) = pk[ch];
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_full[s]);
                TLS(12);
                mbits = nm; rstd = nr;
            }
        } else {
            float rstd = 1.f;
            if (mine > 0) rstd = a.ws.stats[size_t(first) * TM + row].y;
            const int one_ch = (C >= c0 && C < c0 + HALF_COLS) ? (C - c0) >> 3 : -1;                   // warp-uniform
            for (int64_t k = 0; k < mine; ++k) {
                float nr = 1.f;
                if (k + 1 < mine) nr = a.ws.stats[size_t(first + (k + 1) * G) * TM + row].y;
                const int b = int(k & 1);
                TLS(8);
                mbar_wait_relaxed(&acc_full[b], uint32_t(k >> 1) & 1u, 53);
                TLS(9);
                tc_fence_after_sync();
                float v[40];
                ld40(uint32_t(b) * CP, v);
                uint4 pk[HALF_CH];
#pragma unroll
                for (int ch = 0; ch < HALF_CH; ++ch) {
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) w[e] = lrelu_pack(v[8 * ch + 2 * e] * rstd, v[8 * ch + 2 * e + 1] * rstd);
                    if (ch == one_ch) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int col = c0 + 8 * ch + 2 * e;
                            if (col == C) w[e] = 0x3F80u;                                  // u[:, C] = 1: the bias row of dW_ih
                            else if (col + 1 == C) w[e] = (w[e] & 0xFFFFu) | 0x3F800000u;
                            else if (col > C) w[e] = 0u;
                        }
                    }
                    pk[ch] = make_uint4(w[0], w[1], w[2], w[3]);
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[b]);
                TLS(10);
                if (k > 0) mbar_wait_relaxed(&t_empty[0], uint32_t(k - 1) & 1u, 54);                  // dW(k-1) has read the u tile
                TLS(11);
#pragma unroll
                for (int ch = 0; ch < HALF_CH; ++ch) *reinterpret_cast<uint4*>(sT + tile_off(TM, row, HALF_CH * part + ch)) = pk[ch];
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_full[0]);
                TLS(12);
                rstd = nr;
            }
        }
        if (mine > 0) {
            mbar_wait_relaxed(fin, 0, 55);
            tc_fence_after_sync();
            if (roleA) {
                // Q^T: lane = i (second block: i = 128 + lane, lanes 0..31), my 40 columns o  ->  q[o][i] (raw sums: tc_post folds column C+1)
                for (int blk = 0; blk < 2; ++blk) {
                    const int i = blk * 128 + row;
#pragma unroll 1
                    for (int ch = 0; ch < HALF_CH; ++ch) {
                        float d[8];
                        tmem_ld8(tmem_addr(tmem, lane_base, (blk == 0 ? 160u : 320u) + c0 + ch * 8), d);
                        if (i < CP) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const int o = c0 + ch * 8 + e;
                                if (o < C) atomicAdd(a.ws.q + size_t(o) * CP + i, d[e]);
                            }
                        }
                    }
                }
            } else {
                for (int blk = 0; blk < 2; ++blk) {
                    float d[16];
                    tmem_ld16(tmem_addr(tmem, lane_base, (blk == 0 ? 320u : 384u) + 16 * part), d);
                    const int c = blk * 128 + row;
                    if (c < CP) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) if (16 * part + e < NC) atomicAdd(a.ws.dwih + size_t(16 * part + e) * CP + c, d[e]);
                    }
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == TS_W_MMA) tmem_dealloc<512>(tmem);
#undef TLS
    if (a.timeline && int(blockIdx.x) == a.timeline - 1 && tid == 0) {     // FVAE_TIMELINE=1: role A (CTA 0), =2: role B (CTA 1)
        __threadfence();
        const unsigned long long* t = reinterpret_cast<const unsigned long long*>(a.ws.xh) + size_t(blockIdx.x) * 4096;
        printf(roleA ? "fvae split timeline role A (cycles from loop top): MMA: t_full du_next x_ready Q_issued | EPI: wait acc_full computed t_empty_seen arrived\n"
                     : "fvae split timeline role B (cycles from loop top): MMA: gemm1_next t_full g_full dW_issued | EPI: wait acc_full computed t_empty_seen arrived\n");
        for (int64_t k = 2; k < (mine < 14 ? mine : 14); ++k) {
            const unsigned long long t0 = t[k * 16];
            printf("%c item %2d:", roleA ? 'A' : 'B', int(k));
            for (int e = 1; e < 5; ++e) printf(" %6lld", (long long)(t[k * 16 + e] - t0));
            printf("  |");
            for (int e = 8; e < 13; ++e) printf(" %6lld", (long long)(t[k * 16 + e] - t0));
            printf("  | next +%lld\n", (long long)(t[(k + 1) * 16] - t0));
        }
    }
}
