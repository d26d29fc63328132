// Probe (GPU box only): cycles per tcgen05.mma for the operand layouts the front-backward kernel uses -- which of the
// MN-major (weight-gradient) configurations is slow?   nvcc -arch=sm_100a -I../../factorvae_b200/csrc -o umma_probe umma_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include "tc_sm100.cuh"
using namespace fvae::tc;

__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
    return make_smem_desc(saddr, lbo, sbo) | (uint64_t(type & 7u) << 61);
}
struct Cfg { int a_kind, b_kind, N, a_mn, b_mn; };   // kinds: 0 NONE chunk-major, 2 SW128, 4 SW64
// returns descriptor + per-K16-step increment (in 16-byte units) for an operand of `rows` MN extent
__device__ void make_op(uint32_t base, int kind, int mn_major, int extent, uint64_t& d, uint64_t& step) {
    if (!mn_major) {           // K-major
        if (kind == 0) { d = make_smem_desc(base, uint32_t(extent) * 16, 128); step = (2u * extent * 16) >> 4; }
        else if (kind == 2) { d = desc_sw(base, 16, 1024, 2); step = 32 >> 4; }
        else { d = desc_sw(base, 16, 512, 4); step = 32 >> 4; }
    } else {                   // MN-major, 128 K rows per tile
        if (kind == 0) { d = make_smem_desc(base, 128, 2048); step = 256 >> 4; }
        else if (kind == 2) { d = desc_sw(base, 16384, 1024, 2); step = 2048 >> 4; }
        else { d = desc_sw(base, 8192, 512, 4); step = 1024 >> 4; }
    }
}
__global__ void __launch_bounds__(128, 1) probe(Cfg c, int reps, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 180 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    fence_async_smem(); tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t abase = smem_u32(smem), bbase = smem_u32(smem + 64 * 1024);
        uint64_t ad0, as, bd0, bs;
        make_op(abase, c.a_kind, c.a_mn, 128, ad0, as);
        make_op(bbase, c.b_kind, c.b_mn, c.N, bd0, bs);
        const uint32_t idesc = make_idesc_bf16(128, c.N, c.a_mn != 0, c.b_mn != 0);
        const int ksteps = 8;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            uint64_t ad = ad0, bd = bd0;
            for (int ks = 0; ks < ksteps; ++ks) { mma_bf16_ss(tmem, ad, bd, idesc, 1u); ad += as; bd += bs; }
        }
        long long t1 = clock64();
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before_sync(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}
int main() {
    long long* d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct { const char* name; Cfg c; } cases[] = {
        {"GEMM1 : A sw128 K-major x B none K-major, N=160", {2, 0, 160, 0, 0}},
        {"row   : A none K-major  x B none K-major, N=160", {0, 0, 160, 0, 0}},
        {"row   : A none K-major  x B none K-major, N=64 ", {0, 0, 64, 0, 0}},
        {"QA old: A none MN x B none MN, N=160", {0, 0, 160, 1, 1}},
        {"QA128 : A none MN x B sw128 MN, N=128", {0, 2, 128, 1, 1}},
        {"QA32  : A none MN x B sw64  MN, N=32 ", {0, 4, 32, 1, 1}},
        {"QB0   : A sw128 MN x B none MN, N=32 ", {2, 0, 32, 1, 1}},
        {"QB1   : A sw64 MN(M=128 over-read) x B none MN, N=32", {4, 0, 32, 1, 1}},
        {"QB old: A none MN x B none MN, N=32 ", {0, 0, 32, 1, 1}},
        {"DW    : A none MN x B none MN, N=64 ", {0, 0, 64, 1, 1}},
        {"mixed : A sw128 MN x B sw128 MN, N=128", {2, 2, 128, 1, 1}},
        {"mixed : A none K-major x B none MN, N=160", {0, 0, 160, 0, 1}},
        {"mixed : A none MN x B none K-major, N=160", {0, 0, 160, 1, 0}},
    };
    for (auto& cs : cases) {
        const int reps = 16;
        probe<<<1, 128, 200 * 1024>>>(cs.c, reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%-52s  %s  issue %.1f clk/MMA   complete %.1f clk/MMA   (floor %d)\n", cs.name, cudaGetErrorString(e), h[0] / double(reps * 8), h[1] / double(reps * 8), 128 * cs.c.N / 256);
        if (e != cudaSuccess) break;
    }
    return 0;
}
