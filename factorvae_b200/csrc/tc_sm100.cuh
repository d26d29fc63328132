// Minimal hand-written sm_100a tensor-core layer: tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM),
// TMEM alloc / ld, mbarrier, proxy fences, UMMA shared-memory / instruction descriptors.
//
// Shared-memory operand layout used everywhere in this repo ("chunk-major", SWIZZLE_NONE):
//     a tile of R rows x KC*8 bf16 is stored as  tile[chunk c][row r][8 elements]   (16 B per (c,r))
//     byte offset(r, k) = (k/8) * (R*16) + r*16 + (k%8)*2
// Read as a K-major operand (rows = M or N index, k = reduction index):
//     core matrix = 8 rows x 16 B, contiguous 128 B;  SBO (8-row group stride) = 128 B;
//     LBO (stride between the two 8-element k chunks of one K=16 MMA) = R*16 B.
// Read as an MN-major operand (the 8-element chunks run along M/N, rows = reduction index):
//     SBO (stride between 8-element MN chunks) = R*16 B;  LBO (stride between 8-row k groups) = 128 B.
// So one tile serves both the row GEMMs (K-major) and the weight-gradient GEMMs (MN-major).
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor"
// tables (same fields as CUTLASS cute/arch/mma_sm100_desc.hpp).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace fvae {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- mbarrier -----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a descriptor / phase bug shows up as a trapped kernel (an error code at the C ABI)
// instead of a hung GPU.  2^26 polls is seconds; a healthy wait is microseconds.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) asm volatile("trap;");
    }
}

// same, naming the wait site before the trap (the warp-specialised kernels have a dozen hand-offs: a phase bug must say where)
__device__ __forceinline__ void mbar_wait_site(uint64_t* bar, uint32_t parity, int site) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) {
            printf("fvae: mbarrier wait timed out: site %d block %d thread %d parity %u\n", site, int(blockIdx.x), int(threadIdx.x), parity);
            asm volatile("trap;");
        }
    }
}

// Wait with back-off, for the many-warp roles.  Every failed try_wait is a shared-memory access; sixteen idle epilogue warps
// polling flat out take a measurable share of the shared-memory port away from the UMMA operand reads (measured: the same
// four UMMAs took 1060 cycles with the epilogue warps spinning, 320 alone).  The issuer / producer threads keep the tight loop.
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity, int site) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    do {
        __nanosleep(32);
        if (++spins > (1u << 22)) {
            printf("fvae: mbarrier wait timed out: site %d block %d thread %d parity %u\n", site, int(blockIdx.x), int(threadIdx.x), parity);
            asm volatile("trap;");
        }
    } while (!mbar_try_wait(bar, parity));
}

// ---- bulk async copy (TMA, 1-D): global -> shared, completion counted in bytes on an mbarrier ----------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- fences ---------------------------------------------------------------------------------------
// generic-proxy smem writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one warp, power-of-two columns >= 32) ----------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
    static_assert(kCols == 32 || kCols == 64 || kCols == 128 || kCols == 256 || kCols == 512, "TMEM columns");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

// ---- descriptors ------------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= uint64_t((saddr >> 4) & 0x3FFF);            // [0,14)  start address >> 4
    d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;  // [16,30) leading-dimension byte offset >> 4
    d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;  // [32,46) stride-dimension byte offset >> 4
    d |= uint64_t(1) << 46;                          // [46,48) descriptor version = 1
    return d;                                        // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// instruction descriptor for kind::f16: A,B = bf16, D = fp32
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4)                       // c_format = F32
         | (1u << 7)                       // a_format = BF16
         | (1u << 10)                      // b_format = BF16
         | (uint32_t(a_mn_major) << 15)    // a_major: 0 = K-major, 1 = MN-major
         | (uint32_t(b_mn_major) << 16)    // b_major
         | ((N >> 3) << 17)                // n_dim
         | ((M >> 4) << 24);               // m_dim
}

// ---- MMA issue (one thread): D[tmem] (+)= A[smem] * B[smem]^T -----------------------------------------
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Warp-converged issue: ALL 32 lanes of the issuer warp execute this with identical operands, one elected lane issues.  In a
// `if (lane == 0)` region ptxas must assume divergent values and wraps every UMMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY
// loop (~14 instructions, ~78 cycles per UMMA measured in situ -- more than a 128 x 32 x 16 UMMA takes to execute); converged,
// the descriptors move to uniform registers with plain R2UR.
__device__ __forceinline__ void mma_bf16_ss_w(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit_w(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(smem_u32(bar)) : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: warp w reads lanes 32*(w%4)..+31, thread = one lane (row), 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// issue only (no wait): pair with tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 8 consecutive columns of my lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// several 8-column reads of my lane behind ONE wait (each tmem_ld8 pays its own TMEM round trip; on a latency-bound chain like the
// GRU step that is three round trips per gate block instead of one)
__device__ __forceinline__ void tmem_ld8x3(uint32_t t0, uint32_t t1, uint32_t t2, float (&a)[8], float (&b)[8], float (&c)[8]) {
    uint32_t r[24];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%24];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%25];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%26];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23])
        : "r"(t0), "r"(t1), "r"(t2)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[8 + i]); c[i] = __uint_as_float(r[16 + i]); }
}
__device__ __forceinline__ void tmem_ld8x2(uint32_t t0, uint32_t t1, float (&a)[8], float (&b)[8]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%16];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%17];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(t0), "r"(t1)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[8 + i]); }
}
// hardware tanh (MUFU.TANH), |rel err| ~ 2^-11: used only on the bf16 tensor-core path
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
// Packed fp32 pairs (FFMA2 / FADD2 / FMUL2, sm_100): one issue slot of the fma pipe carries two lanes' worth of arithmetic.  The GRU
// gate epilogues are bound by that pipe (a 3-register FFMA holds it for two cycles per warp), and neighbouring hidden units are
// independent, so their pointwise math is written on pairs.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 splat2(float v) { return make_float2(v, v); }
__device__ __forceinline__ float2 tanh2(float2 x) { return make_float2(tanh_fast(x.x), tanh_fast(x.y)); }
__device__ __forceinline__ float2 sigmoid2(float2 x) { return fma2(tanh2(mul2(x, splat2(0.5f))), splat2(0.5f), splat2(0.5f)); }
// 8 bf16 (one 16-byte chunk) <-> 8 floats
__device__ __forceinline__ void unpack8(const uint4& p, float (&v)[8]) {
    const uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
}

// TMEM address: bits [31:16] lane, [15:0] column
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) { return base + (lane << 16) + col; }

// ---- GEMM issue helpers over 128-row chunk-major tiles (one thread issues) -----------------------------------
constexpr uint32_t kTileRows = 128;                 // UMMA M = TMEM lanes
constexpr uint32_t kTileChunk = kTileRows * 16;     // bytes of one 8-column chunk of a 128-row tile

// The descriptors of consecutive k steps differ only in the start-address field (bits [0,14), units of 16 bytes): build
// them once and step with one 64-bit add per operand (shared memory is < 256 KB, so the field never carries out).
// row GEMM: D[128 x N] (tmem column dcol) (+)= A(K-major tile, 128 rows) . B(K-major image, brows rows)^T over k16 steps
__device__ __forceinline__ void issue_row_gemm_acc(uint32_t tmem, uint32_t dcol, uint32_t a_addr, uint32_t b_addr, uint32_t brows,
                                                   uint32_t N, int k16, bool accumulate) {
    const uint32_t idesc = make_idesc_bf16(kTileRows, N, false, false);
    uint64_t ad = make_smem_desc(a_addr, kTileChunk, 128);
    uint64_t bd = make_smem_desc(b_addr, brows * 16, 128);
    const uint64_t astep = (2 * kTileChunk) >> 4, bstep = (2 * brows * 16) >> 4;
    for (int ks = 0; ks < k16; ++ks) {
        mma_bf16_ss(tmem + dcol, ad, bd, idesc, (accumulate || ks > 0) ? 1u : 0u);
        ad += astep; bd += bstep;
    }
}
// weight-gradient GEMM: D[128 x N] (+)= A^T . B with A, B 128-row tiles read MN-major; A's M block starts at chunk a_chunk0
__device__ __forceinline__ void issue_wgrad_acc(uint32_t tmem, uint32_t dcol, uint32_t a_addr, uint32_t a_chunk0, uint32_t b_addr,
                                                uint32_t N, bool accumulate) {
    const uint32_t idesc = make_idesc_bf16(kTileRows, N, true, true);
    uint64_t ad = make_smem_desc(a_addr + a_chunk0 * kTileChunk, 128, kTileChunk);
    uint64_t bd = make_smem_desc(b_addr, 128, kTileChunk);
#pragma unroll
    for (int ks = 0; ks < int(kTileRows) / 16; ++ks) {
        mma_bf16_ss(tmem + dcol, ad, bd, idesc, (accumulate || ks > 0) ? 1u : 0u);
        ad += 256 >> 4; bd += 256 >> 4;
    }
}

// warp-converged forms of the two issue loops (see mma_bf16_ss_w)
__device__ __forceinline__ void issue_row_gemm_w(uint32_t tmem, uint32_t dcol, uint32_t a_addr, uint32_t b_addr, uint32_t brows,
                                                 uint32_t N, int k16) {
    const uint32_t idesc = make_idesc_bf16(kTileRows, N, false, false);
    uint64_t ad = make_smem_desc(a_addr, kTileChunk, 128);
    uint64_t bd = make_smem_desc(b_addr, brows * 16, 128);
    const uint64_t astep = (2 * kTileChunk) >> 4, bstep = (2 * brows * 16) >> 4;
    for (int ks = 0; ks < k16; ++ks) {
        mma_bf16_ss_w(tmem + dcol, ad, bd, idesc, ks > 0 ? 1u : 0u);
        ad += astep; bd += bstep;
    }
}

__device__ __forceinline__ void issue_row_gemm_acc_w(uint32_t tmem, uint32_t dcol, uint32_t a_addr, uint32_t b_addr, uint32_t brows,
                                                     uint32_t N, int k16, bool accumulate) {
    const uint32_t idesc = make_idesc_bf16(kTileRows, N, false, false);
    uint64_t ad = make_smem_desc(a_addr, kTileChunk, 128);
    uint64_t bd = make_smem_desc(b_addr, brows * 16, 128);
    const uint64_t astep = (2 * kTileChunk) >> 4, bstep = (2 * brows * 16) >> 4;
    for (int ks = 0; ks < k16; ++ks) {
        mma_bf16_ss_w(tmem + dcol, ad, bd, idesc, (accumulate || ks > 0) ? 1u : 0u);
        ad += astep; bd += bstep;
    }
}
__device__ __forceinline__ void issue_wgrad_acc_w(uint32_t tmem, uint32_t dcol, uint32_t a_addr, uint32_t a_chunk0, uint32_t b_addr,
                                                  uint32_t N, bool accumulate) {
    const uint32_t idesc = make_idesc_bf16(kTileRows, N, true, true);
    uint64_t ad = make_smem_desc(a_addr + a_chunk0 * kTileChunk, 128, kTileChunk);
    uint64_t bd = make_smem_desc(b_addr, 128, kTileChunk);
#pragma unroll
    for (int ks = 0; ks < int(kTileRows) / 16; ++ks) {
        mma_bf16_ss_w(tmem + dcol, ad, bd, idesc, (accumulate || ks > 0) ? 1u : 0u);
        ad += 256 >> 4; bd += 256 >> 4;
    }
}

// ---- packing helpers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// byte offset of (row, 8-element chunk) in a chunk-major tile of R rows
__device__ __forceinline__ uint32_t tile_off(uint32_t R, uint32_t row, uint32_t chunk) { return chunk * (R * 16u) + row * 16u; }

}  // namespace tc
}  // namespace fvae
