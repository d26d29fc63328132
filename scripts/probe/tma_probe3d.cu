// Probe (GPU box only): the 3-D tensor map of the TMA front kernels -- panel x[S][T][160] bf16 (158 used), box {64|32, 1, 128}
// at (0|64|128, t, tile*128): completes, lands where xs_chunk_off() says, columns >= 158 and rows >= S arrive as zeros.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__global__ void probe(const __grid_constant__ CUtensorMap m128, const __grid_constant__ CUtensorMap m64, int t, int r0, unsigned char* out, int* status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(40960) : "memory");
        const int offs[3] = {0, 16384, 32768}, cs[3] = {0, 64, 128};
        for (int i = 0; i < 3; ++i)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(smem + offs[i])), "l"(i < 2 ? &m128 : &m64), "r"(cs[i]), "r"(t), "r"(r0), "r"(smem_u32(&bar)) : "memory");
        int ok = 0;
        for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
            uint32_t r;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(r) : "r"(smem_u32(&bar)) : "memory");
            ok = r;
        }
        *status = ok;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 40960; i += blockDim.x) out[i] = smem[i];
}
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static uint32_t off(uint32_t r, uint32_t j) {
    if (j < 16) return (j >> 3) * 16384u + r * 128u + (((j & 7u) ^ (r & 7u)) << 4);
    return 32768u + r * 64u + ((((j - 16u) & 3u) ^ ((r >> 1) & 3u)) << 4);
}
int main() {
    const int S = 300, T = 20, C = 158, P = 160;
    std::vector<__nv_bfloat16> h(size_t(S) * T * P);
    for (int s = 0; s < S; ++s) for (int t = 0; t < T; ++t) for (int c = 0; c < P; ++c) h[(size_t(s) * T + t) * P + c] = __float2bfloat16(c < C ? float((s * 7 + t * 3 + c) % 251) : 99.f);
    __nv_bfloat16* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    PFN enc = (PFN)p;
    CUtensorMap m128, m64;
    cuuint64_t dims[3] = {cuuint64_t(C), cuuint64_t(T), cuuint64_t(S)}; cuuint64_t str[2] = {cuuint64_t(P) * 2, cuuint64_t(T) * P * 2}; cuuint32_t es[3] = {1, 1, 1};
    cuuint32_t b128[3] = {64, 1, 128}, b64[3] = {32, 1, 128};
    CUresult r1 = enc(&m128, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, str, b128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&m64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, str, b64, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d %d\n", int(r1), int(r2));
    unsigned char* dout; int* dst; cudaMalloc(&dout, 40960); cudaMalloc(&dst, 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 41984);
    std::vector<unsigned char> o(40960);
    for (int t : {0, 1, 19}) for (int r0 : {0, 256}) {
        cudaMemset(dst, 0xff, 4);
        probe<<<1, 128, 40960, 0>>>(m128, m64, t, r0, dout, dst);
        cudaError_t e = cudaDeviceSynchronize();
        int st; cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost); cudaMemcpy(o.data(), dout, 40960, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int r = 0; r < 128; ++r) for (int c = 0; c < 160; ++c) {
            const float got = __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&o[off(r, c / 8) + (c % 8) * 2]));
            const int s = r0 + r;
            const float want = (s < S && c < C) ? float((s * 7 + t * 3 + c) % 251) : 0.f;
            if (got != want) ++bad;
        }
        printf("t=%d r0=%d: err=%s completed=%d mismatches=%ld\n", t, r0, cudaGetErrorString(e), st, bad);
    }
    return 0;
}
