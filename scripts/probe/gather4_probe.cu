// Probe (GPU box only): cp.async.bulk.tensor.2d tile::gather4 from a row table [R][160] bf16 (pitch 320 B) into a 128-byte-
// swizzled stage: which box shape does the tensor map need, where do the 4 rows land?
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void probe(const __grid_constant__ CUtensorMap m, int c0, int bytes_per_instr, const int* rows, unsigned char* out, int* status) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) smem[i] = 0xEE;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes_per_instr * 2) : "memory");
        for (int g = 0; g < 2; ++g)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                         ::"r"(smem_u32(smem + g * 512)), "l"(&m), "r"(c0), "r"(rows[4 * g]), "r"(rows[4 * g + 1]), "r"(rows[4 * g + 2]), "r"(rows[4 * g + 3]),
                           "r"(smem_u32(&bar)) : "memory");
        int ok = 0;
        for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
            uint32_t r;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(r) : "r"(smem_u32(&bar)) : "memory");
            ok = r;
        }
        *status = ok;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) out[i] = smem[i];
}

typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 0;          // one variant per process: a fault is sticky
    const int R = 1000, P = 160, C = 158;
    std::vector<__nv_bfloat16> h(size_t(R) * P);
    for (int r = 0; r < R; ++r) for (int c = 0; c < P; ++c) h[size_t(r) * P + c] = __float2bfloat16(float((r * 3 + c) % 253));
    __nv_bfloat16* d; cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    PFN enc = (PFN)p;
    const int hrows[8] = {5, 999, 17, 300, 301, 2, 640, 77};
    int* drows; cudaMalloc(&drows, 32); cudaMemcpy(drows, hrows, 32, cudaMemcpyHostToDevice);
    unsigned char* dout; int* dst; cudaMalloc(&dout, 16384); cudaMalloc(&dst, 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 17408);
    const cuuint32_t boxes[3][2] = {{64, 1}, {64, 4}, {64, 8}};
    CUtensorMap m;
    cuuint64_t dims[2] = {cuuint64_t(C), cuuint64_t(R)}; cuuint64_t str[1] = {cuuint64_t(P) * 2}; cuuint32_t es[2] = {1, 1};
    cuuint32_t box[2] = {boxes[which][0], boxes[which][1]};
    CUresult r1 = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box {%u,%u}: encode %d\n", box[0], box[1], int(r1));
    if (r1 != CUDA_SUCCESS) return 0;
    for (int c0 : {0, 64, 128}) {
        cudaMemset(dst, 0xff, 4);
        probe<<<1, 128, 16384, 0>>>(m, c0, 512, drows, dout, dst);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<unsigned char> o(16384);
        int st; cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost); cudaMemcpy(o.data(), dout, 16384, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int r = 0; r < 8; ++r) for (int c = 0; c < 64; ++c) {
            const uint32_t off = r * 128 + ((((c / 8) & 7) ^ (r & 7)) << 4) + (c % 8) * 2;
            const float got = __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&o[off]));
            const float want = (c0 + c < C) ? float((hrows[r] * 3 + c0 + c) % 253) : 0.f;
            if (got != want) ++bad;
        }
        printf("  c0=%d: err=%s completed=%d mismatches(8 rows x 64)=%ld  first bytes row0: %02x %02x row4: %02x %02x\n", c0, cudaGetErrorString(e), st, bad, o[0], o[1], o[512], o[513]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
