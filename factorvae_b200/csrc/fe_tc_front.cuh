// Item kernels of the tensor-core FeatureExtractor (included by fe_tc.cu inside namespace fvae::<anon>):
//   K1  tc_front_fwd_kernel    : LayerNorm -> GEMM1 -> LeakyReLU -> GEMM2 -> GI tile      (reference module.py:26-30);
//                                also saves the xhat / u operand tiles and the LeakyReLU' sign bits for backward
//   K4a tc_q_from_tiles_kernel : du = dGI . W_ih, dpre = du * LeakyReLU', Q += dpre^T [xhat | 1]   (streams saved tiles)
//   K4b tc_wih_from_u_kernel   : dWih += dGI^T [u | 1]                                             (streams saved tiles)
// An item is (sequence tile of 128 stocks, time step): 128 panel rows = one UMMA M = the 128 TMEM lanes.
// K1 runs 512 threads per CTA: thread (row = tid & 127, part = tid >> 7) owns 40 of the 160 columns of its row in
// LayerNorm and in every epilogue (warps w, w+4, w+8, w+12 read the same 32 TMEM lanes, different columns);
// 16 resident warps hide the LDS / TMEM-load latencies that 4 warps cannot.
#pragma once

struct ItemArgs {
    const void* x; int64_t seq_pitch, row_pitch;
    const int32_t* row_index; int64_t num_rows;       // resident panel (NULL: dense windows)
    int items32;                                      // item arithmetic in 32 bits (NT * T * T < 2^32)
    uint32_t t_magic;                                 // floor(2^32 / T) + 1
    int S, T, C, H, NC, HP; int64_t NT;
    int prefetch;            // 1: dedicated raw-row stage, next item's rows are fetched during this item's MMAs
    TcWs ws;
};

constexpr int NSPLIT = 4;                    // column parts per row
constexpr int NTH = TM * NSPLIT;             // 512 threads
constexpr int HALF_COLS = CP / NSPLIT;       // 40 columns per thread
constexpr int HALF_CH = KCH / NSPLIT;        // 5 chunks per thread
constexpr int NPW = HALF_COLS / 2 / 4 + 1;   // 16-byte pieces a thread may touch for its bf16 words (6)
constexpr int NPF = HALF_COLS / 4 + 1;       // ... for its fp32 values (11)

__device__ __forceinline__ void copy_image(unsigned char* dst, const void* src, uint32_t bytes) {
    for (uint32_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
}

// row GEMM / weight-gradient GEMM issue loops: tc_sm100.cuh (issue_row_gemm_acc / issue_wgrad_acc)
__device__ __forceinline__ void issue_row_gemm(uint32_t tmem, uint32_t dcol, uint32_t a_addr, uint32_t b_addr, uint32_t brows,
                                               uint32_t N, int k16) {
    issue_row_gemm_acc(tmem, dcol, a_addr, b_addr, brows, N, k16, false);
}
__device__ __forceinline__ void issue_wgrad(uint32_t tmem, uint32_t dcol, uint32_t a_addr, uint32_t a_chunk0, uint32_t b_addr,
                                            uint32_t N, bool accumulate) {
    issue_wgrad_acc(tmem, dcol, a_addr, a_chunk0, b_addr, N, accumulate);
}

// ---- staging of raw panel rows ------------------------------------------------------------------------------
// A row of the panel (C*2 = 316 B in bf16) starts at a 2- or 4-byte aligned address, so it is copied as the
// 16-byte aligned window that covers it: slot = round16(row bytes) + 16 bytes, fetched with cp.async.cg 16 B
// (LDGSTS.128: no register staging, every row of the item in flight at once).  The consumer adds
// (row address & 15) to find its first element.  A window that would run past the end of the panel
// (last rows only) falls back to element copies.  The slot pitch (336 / 656 B) is an odd multiple of 16, so
// thread-per-row LDS.128 reads are bank-conflict free.
constexpr uint32_t STAGE_BYTES = 43008;      // 128 slots x 336 B (bf16 rows) >= 64 slots x 656 B (fp32 rows)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <typename XT>
__device__ __forceinline__ uint32_t slot_bytes(int C) {
    uint32_t s = ((uint32_t(C) * sizeof(XT) + 15u) & ~15u) + 16u;
    if (((s >> 4) & 1u) == 0u) s += 16u;       // odd number of 16-byte pieces
    return s;
}
template <typename XT>
__device__ __forceinline__ const unsigned char* row_ptr(const ItemArgs& a, int64_t s, int t) {
    if (a.row_index) return reinterpret_cast<const unsigned char*>(static_cast<const XT*>(a.x) + int64_t(a.row_index[s * a.T + t]) * a.row_pitch);
    return reinterpret_cast<const unsigned char*>(static_cast<const XT*>(a.x) + s * a.seq_pitch + int64_t(t) * a.row_pitch);
}
// one past the last byte of the panel allocation a 16-byte window may touch
template <typename XT>
__device__ __forceinline__ const unsigned char* panel_end(const ItemArgs& a) {
    if (a.row_index) return reinterpret_cast<const unsigned char*>(static_cast<const XT*>(a.x) + (a.num_rows - 1) * a.row_pitch + a.C);
    return reinterpret_cast<const unsigned char*>(static_cast<const XT*>(a.x) + int64_t(a.S - 1) * a.seq_pitch + int64_t(a.T - 1) * a.row_pitch + a.C);
}
template <typename XT>
struct Rows { static constexpr int PER_PASS = sizeof(XT) == 2 ? TM : TM / 2; };
// item -> (tile, time step).  Items fit 32 bits in every supported workload (the host checks): a 64-bit integer division
// costs ~100 instructions per thread, and this kernel is issue-bound.
__device__ __forceinline__ void split_item(const ItemArgs& a, int64_t item, int64_t& st, int& t) {
    if (a.items32) {
        // multiply-high by magic = floor(2^32 / T) + 1: exact for item * T < 2^32 (the host sets items32 accordingly); the
        // compare-and-fix keeps it right even at the boundary
        uint32_t q = __umulhi(uint32_t(item), a.t_magic);
        uint32_t r = uint32_t(item) - q * uint32_t(a.T);
        if (r >= uint32_t(a.T)) { ++q; r -= uint32_t(a.T); }
        st = q; t = int(r);
    } else { st = item / a.T; t = int(item % a.T); }
}

// rows [r0, r0 + PER_PASS) of item (st, t) -> slots (zero rows beyond S); one warp per row, asynchronous.
// Fast path (every row of the pass exists and no 16-byte window can cross the end of the panel): lane = piece,
// pointer increments only.
// Resident panel: sOff[row] receives the 16-byte misalignment of the row's source (LayerNorm needs it; it would cost a
// dependent index load there); *carry / next_item: this lane's table row for the NEXT item of the CTA is fetched here,
// one item ahead, so the index load never sits in front of the cp.async issue.
template <typename XT, bool IDX>
__device__ __forceinline__ void load_rows_async(const ItemArgs& a, int64_t st, int t, unsigned char* stage, int r0,
                                                unsigned char* sOff = nullptr, int32_t* carry = nullptr, int64_t next_item = -1) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, C = a.C;
    const uint32_t slot = slot_bytes<XT>(C);
    const int npieces = int(slot / 16);
    constexpr int NW = NTH / 32;
    const int64_t s0 = st * TM + r0;
    if (!IDX && s0 + Rows<XT>::PER_PASS + 1 <= a.S && npieces <= 64) {
        const unsigned char* src = row_ptr<XT>(a, s0 + warp, t);
        const int64_t step = a.seq_pitch * int64_t(sizeof(XT)) * NW;
        uint32_t dst = smem_u32(stage) + uint32_t(warp) * slot + uint32_t(lane) * 16u;
#pragma unroll 4
        for (int rr = warp; rr < Rows<XT>::PER_PASS; rr += NW) {
            const unsigned char* a0 = reinterpret_cast<const unsigned char*>(reinterpret_cast<uintptr_t>(src) & ~uintptr_t(15)) + lane * 16;
            if (lane < npieces) cp_async16(dst, a0);
            if (lane + 32 < npieces) cp_async16(dst + 512u, a0 + 512);
            src += step;
            dst += NW * slot;
        }
        return;
    }
    const unsigned char* x_end = panel_end<XT>(a);
    if (IDX) {
        // resident panel: lane k of the warp fetches the table row of the warp's k-th sequence, one shuffle per row after
        constexpr int RPW = Rows<XT>::PER_PASS / NW;
        int32_t myidx = -1;
        if (carry && *carry != INT32_MIN) myidx = *carry;       // fetched while the previous item was being issued
        else {
            const int64_t s = s0 + warp + int64_t(lane) * NW;
            if (lane < RPW && s < a.S) myidx = a.row_index[s * a.T + t];
        }
        // lane k prepares row k (address, alignment, which path), the warp then walks the rows with three shuffles each:
        // the kernel is issue-bound, so the per-row scalar work must not be replicated over 32 lanes
        unsigned long long my_a0 = 0ull;
        uint32_t my_kind = 0u, my_off = 0u;              // kind 0: zero rows (beyond S), 1: cp.async window, 2: plain loads
        if (lane < RPW && myidx >= 0) {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(static_cast<const XT*>(a.x) + int64_t(myidx) * a.row_pitch);
            const unsigned char* a0 = reinterpret_cast<const unsigned char*>(reinterpret_cast<uintptr_t>(src) & ~uintptr_t(15));
            my_a0 = reinterpret_cast<unsigned long long>(a0);
            my_off = uint32_t(src - a0);
            my_kind = (a0 + slot <= x_end && npieces <= 64) ? 1u : 2u;
        }
        if (sOff && lane < RPW) sOff[warp + lane * NW] = static_cast<unsigned char>(my_off);
        const uint32_t dst0 = smem_u32(stage) + uint32_t(warp) * slot + uint32_t(lane) * 16u;
#pragma unroll
        for (int k = 0; k < RPW; ++k) {
            const unsigned char* a0 = reinterpret_cast<const unsigned char*>(__shfl_sync(0xffffffffu, my_a0, k));
            const uint32_t kind = __shfl_sync(0xffffffffu, my_kind, k);
            const uint32_t dst = dst0 + uint32_t(k * NW) * slot;
            if (kind == 1u) {
                if (lane < npieces) cp_async16(dst, a0 + lane * 16);
                if (lane + 32 < npieces) cp_async16(dst + 512u, a0 + lane * 16 + 512);
            } else {                                                // sequence beyond S, or the last table row: plain stores
                unsigned char* d = stage + size_t(warp + k * NW) * slot;
                for (int pc = lane; pc < npieces; pc += 32) *reinterpret_cast<uint4*>(d + pc * 16) = make_uint4(0, 0, 0, 0);
                __syncwarp();
                if (kind == 2u) {
                    const uint32_t off = __shfl_sync(0xffffffffu, my_off, k);
                    XT* d2 = reinterpret_cast<XT*>(d + off);
                    const XT* srcx = reinterpret_cast<const XT*>(a0 + off);
                    for (int c = lane; c < C; c += 32) d2[c] = srcx[c];
                }
            }
        }
        if (carry) {
            int32_t nx = INT32_MIN;
            if (next_item >= 0) {
                nx = -1;
                int64_t nst; int nt;
                split_item(a, next_item, nst, nt);
                const int64_t s = nst * TM + r0 + warp + int64_t(lane) * NW;
                if (lane < RPW && s < a.S) nx = a.row_index[s * a.T + nt];
            }
            *carry = nx;
        }
        return;
    }
    for (int rr = warp; rr < Rows<XT>::PER_PASS; rr += NW) {
        unsigned char* dst = stage + size_t(rr) * slot;
        const int64_t s = s0 + rr;
        if (s < a.S) {
            const unsigned char* src = row_ptr<XT>(a, s, t);
            const unsigned char* a0 = reinterpret_cast<const unsigned char*>(reinterpret_cast<uintptr_t>(src) & ~uintptr_t(15));
            if (a0 + slot <= x_end) {
                for (int pc = lane; pc < npieces; pc += 32) cp_async16(smem_u32(dst + pc * 16), a0 + pc * 16);
            } else {
                XT* d2 = reinterpret_cast<XT*>(dst + (src - a0));
                for (int c = lane; c < C; c += 32) d2[c] = reinterpret_cast<const XT*>(src)[c];
            }
        } else {
            for (int pc = lane; pc < npieces; pc += 32) *reinterpret_cast<uint4*>(dst + pc * 16) = make_uint4(0, 0, 0, 0);
        }
    }
}

// ---- my 40 features of my row, realigned from the slot into registers -------------------------------------------
template <int W>
__device__ __forceinline__ void realign_words(const uint4* pieces, uint32_t (&w)[HALF_COLS / 2]) {
    uint32_t q[4 * NPW];
#pragma unroll
    for (int i = 0; i < NPW; ++i) {
        if (i < NPW - 1 || W > 0) {
            const uint4 v = pieces[i];
            q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
        } else {
            q[4 * i] = q[4 * i + 1] = q[4 * i + 2] = q[4 * i + 3] = 0u;
        }
    }
#pragma unroll
    for (int j = 0; j < HALF_COLS / 2; ++j) w[j] = q[j + W];
}
template <int W>
__device__ __forceinline__ void realign_floats(const uint4* pieces, float (&v)[HALF_COLS]) {
    uint32_t q[4 * NPF];
#pragma unroll
    for (int i = 0; i < NPF; ++i) {
        if (i < NPF - 1 || W > 0) {
            const uint4 u = pieces[i];
            q[4 * i] = u.x; q[4 * i + 1] = u.y; q[4 * i + 2] = u.z; q[4 * i + 3] = u.w;
        } else {
            q[4 * i] = q[4 * i + 1] = q[4 * i + 2] = q[4 * i + 3] = 0u;
        }
    }
#pragma unroll
    for (int j = 0; j < HALF_COLS; ++j) v[j] = __uint_as_float(q[j + W]);
}

// features [40*part, 40*part + 40) of the row stored in `slot` (first element at byte `off`) -> fp32 registers
__device__ __forceinline__ void fetch_half(const __nv_bfloat16*, const unsigned char* slot, uint32_t off, int half, float (&v)[HALF_COLS]) {
    uint32_t w[HALF_COLS / 2];
    if ((off & 3u) == 0u) {
        const uint4* pieces = reinterpret_cast<const uint4*>(slot) + (HALF_COLS / 8) * half;
        switch (off >> 2) {
            case 0: realign_words<0>(pieces, w); break;
            case 1: realign_words<1>(pieces, w); break;
            case 2: realign_words<2>(pieces, w); break;
            default: realign_words<3>(pieces, w); break;
        }
    } else {       // 2-byte aligned rows (odd pitch): element loads
        const unsigned short* e = reinterpret_cast<const unsigned short*>(slot + off) + HALF_COLS * half;
#pragma unroll
        for (int j = 0; j < HALF_COLS / 2; ++j) w[j] = uint32_t(e[2 * j]) | (uint32_t(e[2 * j + 1]) << 16);
    }
#pragma unroll
    for (int j = 0; j < HALF_COLS / 2; ++j) {
        v[2 * j] = __uint_as_float(w[j] << 16);
        v[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
    }
}
__device__ __forceinline__ void fetch_half(const float*, const unsigned char* slot, uint32_t off, int half, float (&v)[HALF_COLS]) {
    const uint4* pieces = reinterpret_cast<const uint4*>(slot) + (HALF_COLS / 4) * half;
    switch ((off >> 2) & 3u) {
        case 0: realign_floats<0>(pieces, v); break;
        case 1: realign_floats<1>(pieces, v); break;
        case 2: realign_floats<2>(pieces, v); break;
        default: realign_floats<3>(pieces, v); break;
    }
}

// LayerNorm of the rows staged for pass `r0` -> xhat bf16 tile, column C := 1.  Called by all threads (one block
// barrier inside).  Statistics are fp32 sums of x and x^2 over exactly C features (var = E[x^2] - mean^2: the
// cancellation only bites when |mean| >> std, where bf16 operands have already lost the signal; the fp32 mode
// keeps the two-pass form).  Only the thread whose 40 columns straddle C pays for masking.
template <typename XT, bool IDX>
__device__ __forceinline__ void layernorm_pass(const ItemArgs& a, int64_t st, int t, const unsigned char* stage, int r0,
                                               unsigned char* tile, float* sStat, unsigned char* gsave = nullptr) {
    const int tid = threadIdx.x, row = tid & (TM - 1), half = tid >> 7, C = a.C;
    const bool active = row >= r0 && row < r0 + Rows<XT>::PER_PASS;
    const int c0 = HALF_COLS * half;
    const bool partial = c0 + HALF_COLS > C;            // warp-uniform
    const bool tail_only = C >= c0 + HALF_COLS - 8;     // only my last chunk holds columns >= C
    float v[HALF_COLS];
    if (active) {
        const int64_t s = st * TM + row;
        uint32_t off = 0u;
        if (s < a.S) off = IDX ? uint32_t(reinterpret_cast<const unsigned char*>(sStat + 2 * NSPLIT * TM)[row - r0])
                                       : uint32_t(reinterpret_cast<uintptr_t>(row_ptr<XT>(a, s, t)) & 15u);
        fetch_half(static_cast<const XT*>(nullptr), stage + size_t(row - r0) * slot_bytes<XT>(C), off, half, v);
        if (partial) {
            if (tail_only) {
#pragma unroll
                for (int j = HALF_COLS - 8; j < HALF_COLS; ++j) if (c0 + j >= C) v[j] = 0.f;
            } else {
#pragma unroll
                for (int j = 0; j < HALF_COLS; ++j) if (c0 + j >= C) v[j] = 0.f;
            }
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < HALF_COLS; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); }
        sStat[half * TM + row] = s1;
        sStat[NSPLIT * TM + half * TM + row] = s2;
    }
    __syncthreads();
    if (active) {
        const float inv_c = 1.f / float(C);
        const float mean = (sStat[row] + sStat[TM + row] + sStat[2 * TM + row] + sStat[3 * TM + row]) * inv_c;
        const float ex2 = (sStat[4 * TM + row] + sStat[5 * TM + row] + sStat[6 * TM + row] + sStat[7 * TM + row]) * inv_c;
        const float rstd = rsqrtf(fmaxf(ex2 - mean * mean, 0.f) + kLnEps);
        const float shift = -mean * rstd;
#pragma unroll
        for (int j = 0; j < HALF_COLS; ++j) v[j] = fmaf(v[j], rstd, shift);
        if (partial) {
            if (tail_only) {
#pragma unroll
                for (int j = HALF_COLS - 8; j < HALF_COLS; ++j) if (c0 + j >= C) v[j] = (c0 + j == C) ? 1.f : 0.f;
            } else {
#pragma unroll
                for (int j = 0; j < HALF_COLS; ++j) if (c0 + j >= C) v[j] = (c0 + j == C) ? 1.f : 0.f;
            }
        }
#pragma unroll
        for (int ch = 0; ch < HALF_CH; ++ch) {
            const uint4 pk = make_uint4(pack_bf16(v[8 * ch], v[8 * ch + 1]), pack_bf16(v[8 * ch + 2], v[8 * ch + 3]),
                                        pack_bf16(v[8 * ch + 4], v[8 * ch + 5]), pack_bf16(v[8 * ch + 6], v[8 * ch + 7]));
            *reinterpret_cast<uint4*>(tile + tile_off(TM, row, HALF_CH * half + ch)) = pk;
            if (gsave) *reinterpret_cast<uint4*>(gsave + tile_off(TM, row, HALF_CH * half + ch)) = pk;    // xhat tile, saved for backward
        }
    }
}

// all loads of one item (bf16: one pass; fp32: the given pass), committed as ONE cp.async group
template <typename XT, bool IDX>
__device__ __forceinline__ void issue_item_loads(const ItemArgs& a, int64_t item, unsigned char* stage, int r0,
                                                 unsigned char* sOff = nullptr, int32_t* carry = nullptr, int64_t next_item = -1) {
    int64_t st; int t;
    split_item(a, item, st, t);
    load_rows_async<XT, IDX>(a, st, t, stage, r0, sOff, carry, next_item);
    cp_async_commit();
}
// my share of a [NCH x 128 x 16 B] operand tile in HBM -> shared memory, as one cp.async group
__device__ __forceinline__ void issue_tile_load(unsigned char* dst, const unsigned char* src, int nch, int row, int part) {
    for (int ch = part; ch < nch; ch += NSPLIT) cp_async16(smem_u32(dst + tile_off(TM, row, ch)), src + tile_off(TM, row, ch));
    cp_async_commit();
}

// stage (unless already issued) + LayerNorm of one item -> xhat tile.  allow_pending: one younger cp.async
// group (a tile prefetch issued after the rows) may still be in flight.
template <typename XT, bool IDX>
__device__ __forceinline__ void stage_and_normalize(const ItemArgs& a, int64_t item, unsigned char* stage, unsigned char* tile,
                                                    float* sStat, bool already_issued, bool allow_pending = false,
                                                    unsigned char* gsave = nullptr) {
    int64_t st; int t;
    split_item(a, item, st, t);
    for (int r0 = 0; r0 < TM; r0 += Rows<XT>::PER_PASS) {
        if (r0 > 0) __syncthreads();
        if (!(already_issued && r0 == 0)) { load_rows_async<XT, IDX>(a, st, t, stage, r0, reinterpret_cast<unsigned char*>(sStat + 2 * NSPLIT * TM)); cp_async_commit(); }
        if (allow_pending && r0 == 0 && already_issued) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();
        layernorm_pass<XT, IDX>(a, st, t, stage, r0, tile, sStat, gsave);
    }
}

// u = LeakyReLU(acc) for my 40 columns -> bf16 tile, column C := 1.  The bias b1f is already in acc: column C of
// the xhat tile is the constant 1 and column C of the W1g image holds b1f.  max(x, 0.01 x) runs on packed bf16x2.
__device__ __forceinline__ uint32_t lrelu_pack(float lo, float hi) {
    const __nv_bfloat162 x = __floats2bfloat162_rn(lo, hi);
    const __nv_bfloat162 y = __hmax2(x, __hmul2(x, __floats2bfloat162_rn(kLeakySlope, kLeakySlope)));
    return *reinterpret_cast<const uint32_t*>(&y);
}
__device__ __forceinline__ void epilogue_u(uint32_t tmem, uint32_t lane_base, int half, int row, int C, unsigned char* tile,
                                           unsigned char* gtile = nullptr, unsigned long long* gmask = nullptr) {
    const int c0 = HALF_COLS * half;
    const int one_ch = (C >= c0 && C < c0 + HALF_COLS) ? (C - c0) >> 3 : -1;       // warp-uniform
    uint32_t bits_lo = 0u, bits_hi = 0u;           // two 32-bit words: one predicated OR-immediate per column
#pragma unroll
    for (int ch = 0; ch < HALF_CH; ++ch) {
        float v[8];
        tmem_ld8(tmem_addr(tmem, lane_base, c0 + ch * 8), v);
        if (gmask) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {          // branch-free: select an immediate, OR it in
                const uint32_t m = (v[e] > 0.f) ? (1u << (((ch & 3) * 8 + e) & 31)) : 0u;
                if (ch < 4) bits_lo |= m; else bits_hi |= m;
            }
        }
        uint32_t w0 = lrelu_pack(v[0], v[1]), w1 = lrelu_pack(v[2], v[3]), w2 = lrelu_pack(v[4], v[5]), w3 = lrelu_pack(v[6], v[7]);
        if (ch == one_ch) {
            const int e = (C - c0) & 7, q = e >> 1;
            const uint32_t one = 0x3F80u << (16 * (e & 1)), keep = 0xFFFFu << (16 * ((e & 1) ^ 1));
            if (q == 0) w0 = (w0 & keep) | one;
            else if (q == 1) w1 = (w1 & keep) | one;
            else if (q == 2) w2 = (w2 & keep) | one;
            else w3 = (w3 & keep) | one;
        }
        const uint4 pk = make_uint4(w0, w1, w2, w3);
        *reinterpret_cast<uint4*>(tile + tile_off(TM, row, HALF_CH * half + ch)) = pk;
        if (gtile) *reinterpret_cast<uint4*>(gtile + tile_off(TM, row, HALF_CH * half + ch)) = pk;     // saved for backward
    }
    if (gmask) gmask[half * TM + row] = (unsigned long long)bits_lo | ((unsigned long long)bits_hi << 32);
}

// ---- K1: front forward ---------------------------------------------------------------------------------------
// IDX: rows come through fvae_panel.row_index (resident panel); PF: dedicated raw-row stage + software pipeline over items.
// Compile-time variants: the kernel is instruction-cache sensitive (each one carries only its own staging code).
template <typename XT, bool IDX, bool PF>
__global__ void __launch_bounds__(NTH, 1) tc_front_fwd_kernel(ItemArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, row = tid & (TM - 1), half = tid >> 7;
    const int C = a.C, NC = a.NC, NCH = NC / 8;
    unsigned char* sW1 = smem;
    unsigned char* sWih = sW1 + W1_BYTES;
    unsigned char* sA1 = sWih + uint32_t(KCH) * NC * 16;
    unsigned char* sA2 = sA1 + A_BYTES;          // u tile; doubles as the raw-row stage when there is no dedicated one
    unsigned char* sStage = PF ? sA2 + A_BYTES : sA2;
    unsigned char* sTail = PF ? sStage + STAGE_BYTES : sA2 + STAGE_BYTES;
    float* sB1 = reinterpret_cast<float*>(sTail);
    float* sBgi = sB1 + CP;
    float* sStat = sBgi + NC;                    // [2][NSPLIT][128]
    unsigned char* sOff = reinterpret_cast<unsigned char*>(sStat + 2 * NSPLIT * TM);     // [128] source misalignment per row
    uint64_t* bars = reinterpret_cast<uint64_t*>(sOff + TM);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

    copy_image(sW1, a.ws.w1g, W1_BYTES);
    copy_image(sWih, a.ws.wih, uint32_t(KCH) * NC * 16);
    for (int i = tid; i < CP; i += NTH) sB1[i] = a.ws.b1f[i];
    for (int i = tid; i < NC; i += NTH) sBgi[i] = a.ws.bgi[i];
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = uint32_t(warp & 3) * 32u;
    const int64_t nitems = a.NT * a.T;
    constexpr bool prefetch = PF && sizeof(XT) == 2;
    // gi = acc + (b_ih [+ b_hr, b_hz]) -> bf16 GI tile (coalesced 16-byte chunks); the column parts split the chunks
    auto epilogue_gi = [&](int64_t item) {
        unsigned char* gout = reinterpret_cast<unsigned char*>(a.ws.gi) + size_t(item) * NCH * TILE_CH;
#pragma unroll 1
        for (int ch = half; ch < NCH; ch += NSPLIT) {
            float v[8];
            tmem_ld8(tmem_addr(tmem, lane_base, 256 + ch * 8), v);       // bias included: u tile column C = 1, image column C = bias
            *reinterpret_cast<uint4*>(gout + tile_off(TM, row, ch)) =
                make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
    };
    auto issue_gemm1 = [&]() {
        if (tid == 0) {
            tc_fence_after_sync();
            issue_row_gemm(tmem, 0, smem_u32(sA1), smem_u32(sW1), CP, CP, KCH / 2);
            mma_commit(&bars[0]);
        }
    };
    auto issue_gemm2 = [&]() {
        if (tid == 0) {
            tc_fence_after_sync();
            issue_row_gemm(tmem, 256, smem_u32(sA2), smem_u32(sWih), NC, NC, KCH / 2);
            mma_commit(&bars[1]);
        }
    };
    // the xhat tile goes to HBM (saved for backward) straight from the registers LayerNorm packs it in
    auto xh_of = [&](int64_t item) { return reinterpret_cast<unsigned char*>(a.ws.xh) + size_t(item) * A_BYTES; };
    auto save_xhat = [&](int64_t item) {
        unsigned char* g = xh_of(item);
#pragma unroll
        for (int ch = 0; ch < HALF_CH; ++ch)
            *reinterpret_cast<uint4*>(g + tile_off(TM, row, HALF_CH * half + ch)) =
                *reinterpret_cast<const uint4*>(sA1 + tile_off(TM, row, HALF_CH * half + ch));
    };
    const int64_t G = gridDim.x;
    uint32_t phase = 0;
    int32_t carry = INT32_MIN;                      // resident panel: my lane's table row of the next item (INT32_MIN: none)
    if (prefetch) {
        // Software pipeline over items (dedicated raw-row stage):  GEMM2(k) runs under LayerNorm(k+1),
        // GEMM1(k+1) under the GI epilogue of k, and the rows of k+2 stream in under both.
        int64_t item = blockIdx.x;
        if (item < nitems) {
            issue_item_loads<XT, IDX>(a, item, sStage, 0, sOff, &carry, item + G < nitems ? item + G : -1);
            stage_and_normalize<XT, IDX>(a, item, sStage, sA1, sStat, true);
            save_xhat(item);
            fence_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            issue_gemm1();
            if (item + G < nitems) issue_item_loads<XT, IDX>(a, item + G, sStage, 0, sOff, &carry, item + 2 * G < nitems ? item + 2 * G : -1);
        }
        for (; item < nitems; item += G, phase ^= 1) {
            mbar_wait(&bars[0], phase);
            tc_fence_after_sync();
            epilogue_u(tmem, lane_base, half, row, C, sA2, a.ws.u ? reinterpret_cast<unsigned char*>(a.ws.u) + size_t(item) * A_BYTES : nullptr,
                       a.ws.mask + size_t(item) * 4 * TM);
            fence_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            issue_gemm2();
            const int64_t nxt = item + G;
            if (nxt < nitems) {
                stage_and_normalize<XT, IDX>(a, nxt, sStage, sA1, sStat, true);
                save_xhat(nxt);       // A1 is free: GEMM1(item) has completed
                fence_async_smem();
                tc_fence_before_sync();
                __syncthreads();
                issue_gemm1();
                if (nxt + G < nitems) issue_item_loads<XT, IDX>(a, nxt + G, sStage, 0, sOff, &carry, nxt + 2 * G < nitems ? nxt + 2 * G : -1);
            }
            mbar_wait(&bars[1], phase);
            tc_fence_after_sync();
            epilogue_gi(item);
            tc_fence_before_sync();
        }
    } else {
        for (int64_t item = blockIdx.x; item < nitems; item += G, phase ^= 1) {
            stage_and_normalize<XT, IDX>(a, item, sStage, sA1, sStat, false);
            save_xhat(item);
            fence_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            issue_gemm1();
            mbar_wait(&bars[0], phase);
            tc_fence_after_sync();
            epilogue_u(tmem, lane_base, half, row, C, sA2, a.ws.u ? reinterpret_cast<unsigned char*>(a.ws.u) + size_t(item) * A_BYTES : nullptr,
                       a.ws.mask + size_t(item) * 4 * TM);
            fence_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            issue_gemm2();
            mbar_wait(&bars[1], phase);
            tc_fence_after_sync();
            epilogue_gi(item);
            tc_fence_before_sync();
            __syncthreads();
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ---- K4b: dWih += dGI^T [u | 1] from the u tiles saved by K1 -- pure streaming: two operand tiles per item arrive by
// cp.async into a 3-deep ring while the weight-gradient MMAs of older items run; bounded by HBM (61 KB per item).
constexpr int WIH_THREADS = 256;
__global__ void __launch_bounds__(WIH_THREADS, 1) tc_wih_from_u_kernel(ItemArgs a, int nstages) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int NC = a.NC, NCH = NC / 8;
    const int MBW = NC > 128 ? 2 : 1;
    const uint32_t g_bytes = uint32_t(NCH) * TILE_CH, stage_bytes = g_bytes + A_BYTES;      // [dGI tile | u tile]
    unsigned char* sRing = smem;                                   // nstages x [dGI | u]; the dGI M-block over-read runs into u
    uint64_t* full = reinterpret_cast<uint64_t*>(sRing + uint32_t(nstages) * stage_bytes);   // [nstages] tiles landed (complete_tx)
    uint64_t* done = full + nstages;                               // [nstages] the UMMAs that read the stage have completed
    uint64_t* fin = done + nstages;                                // everything issued has completed (one phase: the other
                                                                   // threads cannot track the per-stage phases they never waited on)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(fin + 1);
    if (tid == 0) { for (int i = 0; i < 2 * nstages + 1; ++i) mbar_init(&full[i], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const int64_t nitems = a.NT * a.T, G = gridDim.x;
    const int64_t mine = nitems > int64_t(blockIdx.x) ? (nitems - 1 - blockIdx.x) / G + 1 : 0;
    const bool started = mine > 0;
    // One thread streams everything: bulk copies nstages items ahead, one weight-gradient UMMA group per item.
    if (tid == 0 && mine > 0) {
        auto load = [&](int64_t k) {
            const int stg = int(k % nstages);
            const int64_t item = int64_t(blockIdx.x) + k * G;
            unsigned char* dst = sRing + uint32_t(stg) * stage_bytes;
            mbar_expect_tx(&full[stg], stage_bytes);
            bulk_g2s(dst, reinterpret_cast<const unsigned char*>(a.ws.gi) + size_t(item) * g_bytes, g_bytes, &full[stg]);
            bulk_g2s(dst + g_bytes, reinterpret_cast<const unsigned char*>(a.ws.u) + size_t(item) * A_BYTES, A_BYTES, &full[stg]);
        };
        for (int64_t k = 0; k < mine && k < nstages; ++k) load(k);
        for (int64_t k = 0; k < mine; ++k) {
            const int stg = int(k % nstages);
            mbar_wait(&full[stg], uint32_t(k / nstages) & 1u);
            tc_fence_after_sync();
            const uint32_t base = smem_u32(sRing + uint32_t(stg) * stage_bytes);
            for (int mb = 0; mb < MBW; ++mb) issue_wgrad(tmem, uint32_t(mb) * CP, base, 16 * mb, base + g_bytes, CP, k > 0);
            mma_commit(&done[stg]);
            if (k >= 1 && k - 1 + nstages < mine) {                // the stage of item k-1 is free once its UMMAs are done
                const int ps = int((k - 1) % nstages);
                mbar_wait(&done[ps], uint32_t((k - 1) / nstages) & 1u);
                load(k - 1 + nstages);
            }
        }
        mma_commit(fin);
    }
    __syncwarp();
    // drain: the last item's UMMAs (commits are ordered, so its barrier covers everything before)
    if (mine > 0) mbar_wait(fin, 0);
    tc_fence_after_sync();
    if (started) {
        const int row = tid & (TM - 1), half = tid >> 7;          // 2 column halves of 80
        const uint32_t lane_base = uint32_t(warp & 3) * 32u;
        for (int mb = 0; mb < MBW; ++mb) {
            const int orow = mb * 128 + row;
            const bool ok = orow < NC;
            for (int ch = 0; ch < KCH / 2; ++ch) {
                const int n0 = (CP / 2) * half + ch * 8;
                float v[8];
                tmem_ld8(tmem_addr(tmem, lane_base, uint32_t(mb) * CP + n0), v);
                if (ok) {
                    red_add_v4(a.ws.dwih + size_t(orow) * CP + n0, v[0], v[1], v[2], v[3]);
                    red_add_v4(a.ws.dwih + size_t(orow) * CP + n0 + 4, v[4], v[5], v[6], v[7]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ---- K4a (streaming form): Q += dpre^T [xhat | 1] from the xhat tiles and mask bits saved by K1 ---------------------------
// Per item two operand tiles arrive by cp.async into a ring (xhat 40 KB, dGI 20 KB); tensor work: du = dGI . W_ih,
// then the three Q blocks; CUDA-core work: only the dpre epilogue.  MMA du(k+1) is issued right behind Q(k), so the
// tensor pipe runs while the epilogue threads wait.
template <int NSTG>
__global__ void __launch_bounds__(NTH, 1) tc_q_from_tiles_kernel(ItemArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, row = tid & (TM - 1), half = tid >> 7;
    const int NC = a.NC, NCH = NC / 8;
    const uint32_t g_bytes = uint32_t(NCH) * TILE_CH;
    // stage = [xhat tile | dGI -> dpre tile]; xhat's M-block over-read runs into the dGI tile, which is the larger of
    // the dGI tile (NC columns) and the dpre tile (160 columns)
    const uint32_t d_bytes = g_bytes > A_BYTES ? g_bytes : A_BYTES;
    const uint32_t STG = A_BYTES + d_bytes;
    unsigned char* sWihT = smem;
    unsigned char* sRing = sWihT + uint32_t(NCH) * CP * 16;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + NSTG * STG);          // [0]: du, [1 + stage]: Q of that stage
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + NSTG);
    copy_image(sWihT, a.ws.wihT, uint32_t(NCH) * CP * 16);
    for (uint32_t i = tid; i < NSTG * STG / 16; i += NTH) reinterpret_cast<uint4*>(sRing)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { for (int i = 0; i < 1 + NSTG; ++i) mbar_init(&bars[i], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = uint32_t(warp & 3) * 32u;
    constexpr uint32_t COL_DU = 0, COL_QA = 160, COL_QB0 = 320, COL_QB1 = 352;
    const int64_t nitems = a.NT * a.T, G = gridDim.x;
    auto issue_loads = [&](int64_t it, int stg) {
        if (it < nitems) {
            unsigned char* dst = sRing + uint32_t(stg) * STG;
            const unsigned char* xs = reinterpret_cast<const unsigned char*>(a.ws.xh) + size_t(it) * A_BYTES;
            const unsigned char* gs = reinterpret_cast<const unsigned char*>(a.ws.gi) + size_t(it) * g_bytes;
            for (uint32_t i = tid; i < A_BYTES / 16; i += NTH) cp_async16(smem_u32(dst + i * 16), xs + i * 16);
            for (uint32_t i = tid; i < g_bytes / 16; i += NTH) cp_async16(smem_u32(dst + A_BYTES + i * 16), gs + i * 16);
        }
        cp_async_commit();
    };
    auto issue_du = [&](int stg) {
        if (tid == 0) {
            tc_fence_after_sync();
            issue_row_gemm(tmem, COL_DU, smem_u32(sRing + uint32_t(stg) * STG + A_BYTES), smem_u32(sWihT), CP, CP, NC / 16);
            mma_commit(&bars[0]);
        }
    };
    uint32_t ph_du = 0, ph_q = 0;          // ph_q: one bit per stage
    bool started = false;
    int k = 0;
    int64_t item = blockIdx.x;
    if (NSTG == 1) {
        // both tiles and the W_ih^T image do not fit twice (NC > 160): plain serial loop
        for (; item < nitems; item += G, ++k) {
            unsigned char* sX = sRing;
            unsigned char* sD = sX + A_BYTES;
            const unsigned long long mbits = a.ws.mask[size_t(item) * 4 * TM + half * TM + row];
            if (k > 0) { mbar_wait(&bars[1], ph_q & 1u); ph_q ^= 1u; }
            issue_loads(item, 0);
            cp_async_wait<0>();
            fence_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            issue_du(0);
            mbar_wait(&bars[0], ph_du);
            ph_du ^= 1;
            tc_fence_after_sync();
#pragma unroll
            for (int ch = 0; ch < HALF_CH; ++ch) {
                float d[8];
                tmem_ld8(tmem_addr(tmem, lane_base, COL_DU + HALF_COLS * half + ch * 8), d);
#pragma unroll
                for (int e = 0; e < 8; ++e) d[e] *= ((mbits >> (ch * 8 + e)) & 1ull) ? 1.f : kLeakySlope;
                *reinterpret_cast<uint4*>(sD + tile_off(TM, row, HALF_CH * half + ch)) =
                    make_uint4(pack_bf16(d[0], d[1]), pack_bf16(d[2], d[3]), pack_bf16(d[4], d[5]), pack_bf16(d[6], d[7]));
            }
            fence_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after_sync();
                issue_wgrad(tmem, COL_QA, smem_u32(sD), 0, smem_u32(sX), CP, started);
                issue_wgrad(tmem, COL_QB0, smem_u32(sX), 0, smem_u32(sD) + 16 * TILE_CH, 32, started);
                issue_wgrad(tmem, COL_QB1, smem_u32(sX), 16, smem_u32(sD) + 16 * TILE_CH, 32, started);
                mma_commit(&bars[1]);
            }
            started = true;
        }
        if (k > 0) mbar_wait(&bars[1], ph_q & 1u);
        k = 0;                              // the ring epilogue below must not wait again
    }
    int64_t next_load = blockIdx.x;
    if (NSTG > 1) for (int sidx = 0; sidx < NSTG - 1; ++sidx) { issue_loads(next_load, sidx); next_load += G; }
    if (NSTG > 1 && item < nitems) {
        cp_async_wait<(NSTG >= 2 ? NSTG - 2 : 0)>();         // loads of item 0
        fence_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        issue_du(0);
    }
    for (; NSTG > 1 && item < nitems; item += G, ++k) {
        const int stg = k % NSTG;
        unsigned char* sX = sRing + uint32_t(stg) * STG;
        unsigned char* sD = sX + A_BYTES;
        const unsigned long long mbits = a.ws.mask[size_t(item) * 4 * TM + half * TM + row];
        // refill the stage item k-1 used (its Q MMAs were committed one iteration ago)
        const int refill = (k + NSTG - 1) % NSTG;
        if (k > 0) { mbar_wait(&bars[1 + refill], (ph_q >> refill) & 1u); ph_q ^= 1u << refill; }
        issue_loads(next_load, refill);
        next_load += G;
        mbar_wait(&bars[0], ph_du);        // du(k) is in TMEM
        ph_du ^= 1;
        tc_fence_after_sync();
#pragma unroll
        for (int ch = 0; ch < HALF_CH; ++ch) {
            float d[8];
            tmem_ld8(tmem_addr(tmem, lane_base, COL_DU + HALF_COLS * half + ch * 8), d);
#pragma unroll
            for (int e = 0; e < 8; ++e) d[e] *= ((mbits >> (ch * 8 + e)) & 1ull) ? 1.f : kLeakySlope;
            *reinterpret_cast<uint4*>(sD + tile_off(TM, row, HALF_CH * half + ch)) =
                make_uint4(pack_bf16(d[0], d[1]), pack_bf16(d[2], d[3]), pack_bf16(d[4], d[5]), pack_bf16(d[6], d[7]));
        }
        const bool has_next = item + G < nitems;
        if (has_next) cp_async_wait<(NSTG >= 2 ? NSTG - 2 : 0)>(); else cp_async_wait<0>();     // loads of item k+1 (needed by du(k+1) below)
        fence_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
            issue_wgrad(tmem, COL_QA, smem_u32(sD), 0, smem_u32(sX), CP, started);                       // rows o < 128
            issue_wgrad(tmem, COL_QB0, smem_u32(sX), 0, smem_u32(sD) + 16 * TILE_CH, 32, started);      // rows o >= 128, i < 128
            issue_wgrad(tmem, COL_QB1, smem_u32(sX), 16, smem_u32(sD) + 16 * TILE_CH, 32, started);     // rows o >= 128, i >= 128
            mma_commit(&bars[1 + stg]);
        }
        started = true;
        if (has_next) issue_du((k + 1) % NSTG);       // queued right behind Q(k): COL_DU was fully read before the barrier above
    }
    if (k > 0) { const int last = (k - 1) % NSTG; mbar_wait(&bars[1 + last], (ph_q >> last) & 1u); }
    cp_async_wait<0>();
    tc_fence_after_sync();
    if (started) {
        const int C = a.C;
        for (int ch = 0; ch < HALF_CH; ++ch) {
            const int n0 = HALF_COLS * half + ch * 8;
            float d[8];
            tmem_ld8(tmem_addr(tmem, lane_base, COL_QA + n0), d);
            if (row < C) {
#pragma unroll
                for (int e = 0; e < 8; ++e) atomicAdd(a.ws.q + size_t(row) * CP + n0 + e, d[e]);
            }
        }
        for (int blk = 0; blk < 2; ++blk) {
            float d[8];
            tmem_ld8(tmem_addr(tmem, lane_base, (blk == 0 ? COL_QB0 : COL_QB1) + 8 * half), d);
            const int i = blk * 128 + row;
            if (i < CP) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int o = 128 + 8 * half + e;
                    if (o < C) atomicAdd(a.ws.q + size_t(o) * CP + i, d[e]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ---- K4a (streaming variant): same math as tc_q_from_tiles_kernel, buffers decoupled so nothing waits on HBM -----
// Shared memory: 2 xhat stages | 3 dGI stages | 1 dpre tile | W_ih^T image.  One thread feeds the stages with 1-D bulk
// copies (complete_tx on mbarriers): dGI three items ahead (it gates du), xhat as soon as the Q MMAs of the item that
// used the stage have completed (it is only needed by Q, one and a half iterations later).  du(k+1) is issued in front
// of Q(k), so the dpre epilogue of item k+1 overlaps the Q MMAs of item k.  Used when the buffers fit (NC <= 96).
constexpr int QS_THREADS = NTH + 32;          // 512 epilogue threads + one warp that feeds the stages and issues the UMMAs
// NDP dpre tiles / NG dGI stages: (2, 2) when it fits -- the epilogue of item k+1 then does not wait for the Q MMAs of item k --
// else (1, 3).
template <int NDP, int NG>
__global__ void __launch_bounds__(QS_THREADS, 1) tc_q_stream_kernel(ItemArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, row = tid & (TM - 1), half = tid >> 7;
    const int NC = a.NC, NCH = NC / 8;
    const uint32_t g_bytes = uint32_t(NCH) * TILE_CH;
    unsigned char* sX = smem;                                   // [2][A_BYTES]; the M-block over-read of stage 1 runs into sG
    unsigned char* sG = sX + 2 * A_BYTES;                       // [NG][g_bytes]
    unsigned char* sD = sG + NG * g_bytes;                      // [NDP] dpre tiles
    unsigned char* sWihT = sD + NDP * A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sWihT + uint32_t(NCH) * CP * 16);
    uint64_t* full_x = bars;            // [2]
    uint64_t* full_g = bars + 2;        // [NG <= 3]
    uint64_t* bar_du = bars + 5;
    uint64_t* bar_q = bars + 6;         // [2]
    uint64_t* dpre_ready = bars + 8;    // all epilogue threads have written their part of the dpre tile (and read du)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    const bool issuer = warp == NTH / 32;
    copy_image(sWihT, a.ws.wihT, uint32_t(NCH) * CP * 16);
    for (uint32_t i = tid; i < (2 * A_BYTES + NG * g_bytes + NDP * A_BYTES) / 16; i += QS_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1); mbar_init(dpre_ready, NTH); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = uint32_t(warp & 3) * 32u;
    constexpr uint32_t COL_DU = 0, COL_QA = 160, COL_QB0 = 320, COL_QB1 = 352;
    const int64_t nitems = a.NT * a.T, G = gridDim.x;
    const int64_t mine = nitems > int64_t(blockIdx.x) ? (nitems - 1 - blockIdx.x) / G + 1 : 0;   // my items: blockIdx.x + k G
    auto load_x = [&](int64_t k) {      // one thread
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.ws.xh) + size_t(blockIdx.x + k * G) * A_BYTES;
        mbar_expect_tx(&full_x[k & 1], A_BYTES);
        bulk_g2s(sX + (k & 1) * A_BYTES, src, A_BYTES, &full_x[k & 1]);
    };
    auto load_g = [&](int64_t k) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.ws.gi) + size_t(blockIdx.x + k * G) * g_bytes;
        mbar_expect_tx(&full_g[k % NG], g_bytes);
        bulk_g2s(sG + (k % NG) * g_bytes, src, g_bytes, &full_g[k % NG]);
    };
    auto issue_du = [&](int64_t k) {    // one thread; waits for the dGI tile of item k
        mbar_wait(&full_g[k % NG], uint32_t(k / NG) & 1u);
        issue_row_gemm(tmem, COL_DU, smem_u32(sG + (k % NG) * g_bytes), smem_u32(sWihT), CP, CP, NC / 16);
        mma_commit(bar_du);
    };
    if (mine > 0 && issuer) {
        if ((tid & 31) == 0) {
            for (int64_t k0 = 0; k0 < NG && k0 < mine; ++k0) load_g(k0);
            load_x(0); if (mine > 1) load_x(1);
            tc_fence_after_sync();
            issue_du(0);
            for (int64_t k = 0; k < mine; ++k) {
                const bool has_next = k + 1 < mine;
                mbar_wait(bar_du, uint32_t(k) & 1u);           // du(k) done: its dGI stage is free
                if (k + NG < mine) load_g(k + NG);
                if (k > 0) {                                   // Q(k-1) done: xhat stage (k-1)&1 is free
                    mbar_wait(&bar_q[(k - 1) & 1], uint32_t((k - 1) >> 1) & 1u);
                    if (has_next) load_x(k + 1);
                }
                mbar_wait(dpre_ready, uint32_t(k) & 1u);       // dpre(k) is in shared memory, du(k) has been read out of TMEM
                tc_fence_after_sync();
                if (has_next) issue_du(k + 1);                 // in front of Q(k)
                mbar_wait(&full_x[k & 1], uint32_t(k >> 1) & 1u);
                const uint32_t xs = smem_u32(sX + (k & 1) * A_BYTES), ds = smem_u32(sD + (NDP == 2 ? (k & 1) : 0) * A_BYTES);
                issue_wgrad(tmem, COL_QA, ds, 0, xs, CP, k > 0);                        // rows o < 128
                issue_wgrad(tmem, COL_QB0, xs, 0, ds + 16 * TILE_CH, 32, k > 0);         // rows o >= 128, i < 128
                issue_wgrad(tmem, COL_QB1, xs, 16, ds + 16 * TILE_CH, 32, k > 0);        // rows o >= 128, i >= 128
                mma_commit(&bar_q[k & 1]);
            }
        }
        __syncwarp();
    } else if (mine > 0) {
        unsigned long long mbits = a.ws.mask[size_t(blockIdx.x) * 4 * TM + half * TM + row];
        for (int64_t k = 0; k < mine; ++k) {
            const bool has_next = k + 1 < mine;
            unsigned long long mnext = 0ull;
            if (has_next) mnext = a.ws.mask[size_t(blockIdx.x + (k + 1) * G) * 4 * TM + half * TM + row];
            mbar_wait(bar_du, uint32_t(k) & 1u);               // du(k) is in TMEM
            tc_fence_after_sync();
            uint4 pk[HALF_CH];
#pragma unroll
            for (int ch = 0; ch < HALF_CH; ++ch) {
                float d[8];
                tmem_ld8(tmem_addr(tmem, lane_base, COL_DU + HALF_COLS * half + ch * 8), d);
#pragma unroll
                for (int e = 0; e < 8; ++e) d[e] *= ((mbits >> (ch * 8 + e)) & 1ull) ? 1.f : kLeakySlope;
                pk[ch] = make_uint4(pack_bf16(d[0], d[1]), pack_bf16(d[2], d[3]), pack_bf16(d[4], d[5]), pack_bf16(d[6], d[7]));
            }
            // my dpre tile is free once the Q MMAs that read it are done: item k-1 with one tile, item k-2 with two
            if (NDP == 1) { if (k > 0) mbar_wait(&bar_q[(k - 1) & 1], uint32_t((k - 1) >> 1) & 1u); }
            else if (k > 1) mbar_wait(&bar_q[k & 1], uint32_t((k - 2) >> 1) & 1u);
            unsigned char* sDk = sD + (NDP == 2 ? (k & 1) : 0) * A_BYTES;
#pragma unroll
            for (int ch = 0; ch < HALF_CH; ++ch) *reinterpret_cast<uint4*>(sDk + tile_off(TM, row, HALF_CH * half + ch)) = pk[ch];
            fence_async_smem();
            tc_fence_before_sync();
            mbar_arrive(dpre_ready);
            mbits = mnext;
        }
        mbar_wait(&bar_q[(mine - 1) & 1], uint32_t((mine - 1) >> 1) & 1u);
        tc_fence_after_sync();
        const int C = a.C;
        for (int ch = 0; ch < HALF_CH; ++ch) {
            const int n0 = HALF_COLS * half + ch * 8;
            float d[8];
            tmem_ld8(tmem_addr(tmem, lane_base, COL_QA + n0), d);
            if (row < C) {
                red_add_v4(a.ws.q + size_t(row) * CP + n0, d[0], d[1], d[2], d[3]);
                red_add_v4(a.ws.q + size_t(row) * CP + n0 + 4, d[4], d[5], d[6], d[7]);
            }
        }
        for (int blk = 0; blk < 2; ++blk) {
            float d[8];
            tmem_ld8(tmem_addr(tmem, lane_base, (blk == 0 ? COL_QB0 : COL_QB1) + 8 * half), d);
            const int i = blk * 128 + row;
            if (i < CP) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int o = 128 + 8 * half + e;
                    if (o < C) atomicAdd(a.ws.q + size_t(o) * CP + i, d[e]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ---- K4b (recompute variant, NC <= 128): dW_ih += dGI^T [u | 1] with u = LeakyReLU(xhat W1g^T + b1f) REBUILT from the saved
// xhat tile, so the forward kernel does not have to write the u tiles at all (40 KB per item, 29 % of K1's HBM traffic;
// K1 is HBM-write bound).  Same bytes in as the streaming variant (xhat instead of u), one extra UMMA group (GEMM1,
// 128x160x160) and the LeakyReLU epilogue per item -- both hidden behind the HBM stream:
//   issuer warp : bulk copies (xhat 2 stages, dGI 2 stages), GEMM1(k+1) -> PRE[(k+1)&1] one item ahead,
//                 weight-gradient UMMAs of item k once the epilogue threads have written u(k)
//   512 threads : PRE[k&1] -> LeakyReLU -> bf16 u tile (identical to K1's epilogue_u: same rounding, ones column at C)
constexpr int WR_THREADS = NTH + 32;
__global__ void __launch_bounds__(WR_THREADS, 1) tc_wih_recompute_kernel(ItemArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, row = tid & (TM - 1), half = tid >> 7;
    const int NC = a.NC, NCH = NC / 8, C = a.C;
    const uint32_t g_bytes = uint32_t(NCH) * TILE_CH;
    unsigned char* sX = smem;                                   // [2][A_BYTES] xhat stages
    unsigned char* sG = sX + 2 * A_BYTES;                       // [2][g_bytes] dGI stages; the M-block over-read runs into what follows
    unsigned char* sU = sG + 2 * g_bytes;                       // u tile
    unsigned char* sW1 = sU + A_BYTES;                          // W1g image (bias column folded)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW1 + W1_BYTES);
    uint64_t* full_x = bars;            // [2]
    uint64_t* full_g = bars + 2;        // [2]
    uint64_t* bar_pre = bars + 4;       // [2] GEMM1 into PRE[i] complete
    uint64_t* bar_w = bars + 6;         // weight-gradient UMMAs of the last issued item complete
    uint64_t* u_ready = bars + 7;       // all epilogue threads wrote their part of u (and read PRE)
    uint64_t* fin = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    const bool issuer = warp == NTH / 32;
    copy_image(sW1, a.ws.w1g, W1_BYTES);
    for (uint32_t i = tid; i < (2 * A_BYTES + 2 * g_bytes + A_BYTES) / 16; i += WR_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { for (int i = 0; i < 9; ++i) mbar_init(&bars[i], 1); mbar_init(u_ready, NTH); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = uint32_t(warp & 3) * 32u;
    constexpr uint32_t COL_PRE = 0, COL_DW = 320;              // PRE[0] [0,160) | PRE[1] [160,320) | dW_ih [320,480)
    const int64_t nitems = a.NT * a.T, G = gridDim.x;
    const int64_t mine = nitems > int64_t(blockIdx.x) ? (nitems - 1 - blockIdx.x) / G + 1 : 0;
    if (mine > 0 && issuer) {
        if ((tid & 31) == 0) {
            auto load_x = [&](int64_t k) {
                mbar_expect_tx(&full_x[k & 1], A_BYTES);
                bulk_g2s(sX + (k & 1) * A_BYTES, reinterpret_cast<const unsigned char*>(a.ws.xh) + size_t(blockIdx.x + k * G) * A_BYTES, A_BYTES, &full_x[k & 1]);
            };
            auto load_g = [&](int64_t k) {
                mbar_expect_tx(&full_g[k & 1], g_bytes);
                bulk_g2s(sG + (k & 1) * g_bytes, reinterpret_cast<const unsigned char*>(a.ws.gi) + size_t(blockIdx.x + k * G) * g_bytes, g_bytes, &full_g[k & 1]);
            };
            auto issue_pre = [&](int64_t k) {                   // GEMM1 of item k into PRE[k & 1]
                mbar_wait(&full_x[k & 1], uint32_t(k >> 1) & 1u);
                tc_fence_after_sync();
                issue_row_gemm(tmem, COL_PRE + uint32_t(k & 1) * CP, smem_u32(sX + (k & 1) * A_BYTES), smem_u32(sW1), CP, CP, KCH / 2);
                mma_commit(&bar_pre[k & 1]);
            };
            load_x(0); load_g(0);
            if (mine > 1) { load_x(1); load_g(1); }
            issue_pre(0);
            for (int64_t k = 0; k < mine; ++k) {
                // PRE[(k+1)&1] was last read by the epilogue of item k-1, which has signalled u_ready(k-1) (waited below, last turn)
                if (k + 1 < mine) issue_pre(k + 1);
                if (k + 2 < mine) {                             // xhat stage k&1 is free as soon as GEMM1(k) has completed
                    mbar_wait(&bar_pre[k & 1], uint32_t(k >> 1) & 1u);
                    load_x(k + 2);
                }
                mbar_wait(u_ready, uint32_t(k) & 1u);           // u(k) is in shared memory, PRE[k&1] has been read
                mbar_wait(&full_g[k & 1], uint32_t(k >> 1) & 1u);
                tc_fence_after_sync();
                issue_wgrad(tmem, COL_DW, smem_u32(sG + (k & 1) * g_bytes), 0, smem_u32(sU), CP, k > 0);
                mma_commit(bar_w);
                if (k + 2 < mine) {                             // dGI stage k&1 is free once these UMMAs are done
                    mbar_wait(bar_w, uint32_t(k) & 1u);
                    load_g(k + 2);
                }
            }
            mma_commit(fin);
        }
        __syncwarp();
    } else if (mine > 0) {
        for (int64_t k = 0; k < mine; ++k) {
            mbar_wait(&bar_pre[k & 1], uint32_t(k >> 1) & 1u);
            tc_fence_after_sync();
            uint4 pk[HALF_CH];
            const int c0 = HALF_COLS * half;
            const int one_ch = (C >= c0 && C < c0 + HALF_COLS) ? (C - c0) >> 3 : -1;
#pragma unroll
            for (int ch = 0; ch < HALF_CH; ++ch) {
                float v[8];
                tmem_ld8(tmem_addr(tmem, lane_base, COL_PRE + uint32_t(k & 1) * CP + c0 + ch * 8), v);
                uint32_t w0 = lrelu_pack(v[0], v[1]), w1 = lrelu_pack(v[2], v[3]), w2 = lrelu_pack(v[4], v[5]), w3 = lrelu_pack(v[6], v[7]);
                if (ch == one_ch) {
                    const int e = (C - c0) & 7, q = e >> 1;
                    const uint32_t one = 0x3F80u << (16 * (e & 1)), keep = 0xFFFFu << (16 * ((e & 1) ^ 1));
                    if (q == 0) w0 = (w0 & keep) | one;
                    else if (q == 1) w1 = (w1 & keep) | one;
                    else if (q == 2) w2 = (w2 & keep) | one;
                    else w3 = (w3 & keep) | one;
                }
                pk[ch] = make_uint4(w0, w1, w2, w3);
            }
            if (k > 0) mbar_wait(bar_w, uint32_t(k - 1) & 1u);  // the UMMAs that read u(k-1) are done: the u tile is free
#pragma unroll
            for (int ch = 0; ch < HALF_CH; ++ch) *reinterpret_cast<uint4*>(sU + tile_off(TM, row, HALF_CH * half + ch)) = pk[ch];
            fence_async_smem();
            tc_fence_before_sync();
            mbar_arrive(u_ready);
        }
        mbar_wait(fin, 0);
        tc_fence_after_sync();
        // flush dW_ih: lane = permuted gate row, my 40 columns
#pragma unroll
        for (int ch = 0; ch < HALF_CH; ++ch) {                  // the TMEM load is warp-collective: every lane issues it
            const int n0 = HALF_COLS * half + ch * 8;
            float v[8];
            tmem_ld8(tmem_addr(tmem, lane_base, COL_DW + n0), v);
            if (row < NC) {
                red_add_v4(a.ws.dwih + size_t(row) * CP + n0, v[0], v[1], v[2], v[3]);
                red_add_v4(a.ws.dwih + size_t(row) * CP + n0 + 4, v[4], v[5], v[6], v[7]);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}
