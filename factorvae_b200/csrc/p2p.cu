// One-shot all-reduce of the flat gradient buffer over NVLink peer memory (the step's single
// collective).  The reference has no counterpart (no data parallelism at all); the baseline is one ncclAllReduce, which
// for 0.25 - 2.2 MB is pure latency (~45 us measured on 8 x B200 behind a 1.07 ms step).  Here every rank owns a
// communication buffer that its peers map through CUDA IPC:
//     data  [2 parities][world][n_pad] fp32   slot (parity, r) receives rank r's vector
//     flags [2 parities][world][kMaxChunks]    slot flag = epoch once chunk c of rank r's vector has landed
// One kernel per all-reduce, CTA = chunk: (1) push my chunk into slot (parity, my rank) of EVERY peer with 16-byte stores
// over NVLink, (2) system-scope fence, publish the chunk's flag at every peer, (3) wait until all `world` flags of this chunk
// in MY buffer carry the epoch, (4) sum the `world` slots in rank order (every rank adds in the same order: bit-identical
// results everywhere) and scale.  CTAs never wait on each other, only on the same chunk of the peers; the grid is far below
// the SM count, so all CTAs are resident and the spin cannot deadlock.  Slots alternate with the epoch's parity: a rank can
// be at most one all-reduce ahead of a peer (it needs the peer's contribution to finish its own), so parity p is never
// rewritten while a peer still sums it.
#include <cstdio>
#include <cstring>

#include "fvae_common.cuh"

namespace fvae {
namespace {

constexpr int kMaxChunks = 64;
constexpr int kP2PThreads = 256;

struct P2PLayout { int64_t n_pad; int world; };
__host__ __device__ inline int64_t p2p_data_bytes(int64_t n_pad, int world) { return int64_t(2) * world * n_pad * 4; }
__host__ __device__ inline int64_t p2p_total_bytes(int64_t n_pad, int world) { return p2p_data_bytes(n_pad, world) + int64_t(2) * world * kMaxChunks * 4; }

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kP2PThreads) p2p_allreduce_kernel(float* __restrict__ inout, int64_t n, char* const* __restrict__ peers,
                                                                    int world, int rank, uint32_t epoch, float scale, int64_t n_pad, int nch) {
    const int c = blockIdx.x, tid = threadIdx.x;
    const int par = int(epoch & 1u);
    const int64_t chunk = ((n + nch - 1) / nch + 3) & ~int64_t(3);
    const int64_t lo = int64_t(c) * chunk, hi = lo + chunk < n ? lo + chunk : n;
    const int64_t n4 = lo < hi ? (hi - lo + 3) / 4 : 0;                     // 16-byte pieces (n_pad is a multiple of 4: the tail is padding)
    const float4* src = reinterpret_cast<const float4*>(inout + lo);
    for (int p = 0; p < world; ++p) {
        const int q = (rank + p) % world;                                    // stagger the targets over the ranks
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(peers[q]) + (int64_t(par) * world + rank) * n_pad + lo);
        for (int64_t i = tid; i < n4; i += kP2PThreads) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) {
        uint32_t* f = reinterpret_cast<uint32_t*>(peers[tid] + p2p_data_bytes(n_pad, world)) + (int64_t(par) * world + rank) * kMaxChunks + c;
        st_release_sys(f, epoch);
    }
    if (tid < world) {
        const uint32_t* f = reinterpret_cast<const uint32_t*>(peers[rank] + p2p_data_bytes(n_pad, world)) + (int64_t(par) * world + tid) * kMaxChunks + c;
        unsigned long long spins = 0;
        while (ld_acquire_sys(f) != epoch) {
            if (++spins > (1ull << 31)) { printf("fvae: p2p all-reduce timed out waiting for rank %d (chunk %d, epoch %u)\n", tid, c, epoch); asm volatile("trap;"); }
        }
    }
    __syncthreads();
    const float4* mine = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(peers[rank]) + int64_t(par) * world * n_pad + lo);
    float4* out = reinterpret_cast<float4*>(inout + lo);
    const int64_t slot4 = n_pad / 4;
    for (int64_t i = tid; i < n4; i += kP2PThreads) {
        float4 acc = __ldcv(mine + i);                                        // peers wrote these lines: bypass any stale L1 copy
        for (int r = 1; r < world; ++r) {
            const float4 v = __ldcv(mine + int64_t(r) * slot4 + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
        out[i] = acc;
    }
}

}  // namespace
}  // namespace fvae

using namespace fvae;

extern "C" {

int64_t fvae_p2p_buffer_bytes(int64_t n, int32_t world) {
    if (n <= 0 || world <= 0) return FVAE_ERR_SHAPE;
    return p2p_total_bytes((n + 3) & ~int64_t(3), world);
}

// cudaMalloc (not a pooled allocation: the block must be exportable) + zero + IPC handle (64 bytes)
int fvae_p2p_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle64) {
    if (!dev_ptr || !handle64) return FVAE_ERR_NULL;
    if (bytes <= 0) return FVAE_ERR_SHAPE;
    cudaError_t e = cudaMalloc(dev_ptr, size_t(bytes));
    if (e != cudaSuccess) return int(e);
    if ((e = cudaMemset(*dev_ptr, 0, size_t(bytes))) != cudaSuccess) return int(e);
    cudaIpcMemHandle_t h;
    if ((e = cudaIpcGetMemHandle(&h, *dev_ptr)) != cudaSuccess) return int(e);
    static_assert(sizeof(h) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    return int(cudaDeviceSynchronize());
}
int fvae_p2p_open(const unsigned char* handle64, void** dev_ptr) {
    if (!dev_ptr || !handle64) return FVAE_ERR_NULL;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    return int(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
}
int fvae_p2p_close(void* dev_ptr) { return dev_ptr ? int(cudaIpcCloseMemHandle(dev_ptr)) : FVAE_ERR_NULL; }
int fvae_p2p_free(void* dev_ptr) { return dev_ptr ? int(cudaFree(dev_ptr)) : FVAE_ERR_NULL; }

// inout[n] <- scale * sum over ranks of inout[n]; peer_bases: DEVICE array [world] of the ranks' buffer base pointers as mapped
// in this process (own buffer at [rank]); epoch: 1, 2, 3, ... identical on all ranks; n must fit the buffers (n <= n at alloc).
int fvae_p2p_allreduce(float* inout, int64_t n, void* const* peer_bases, int32_t world, int32_t rank, uint32_t epoch, float scale,
                       int64_t n_alloc, void* stream) {
    if (!inout || !peer_bases) return FVAE_ERR_NULL;
    if (n <= 0 || n > n_alloc || world <= 0 || world > kP2PThreads || rank < 0 || rank >= world || epoch == 0) return FVAE_ERR_SHAPE;
    if (reinterpret_cast<uintptr_t>(inout) % 16 != 0) return FVAE_ERR_WORKSPACE;
    const int64_t n_pad = (n_alloc + 3) & ~int64_t(3);
    int nch = int((n + 4095) / 4096);                     // ~16 KB of fp32 per CTA
    if (nch > kMaxChunks) nch = kMaxChunks;
    if (nch < 1) nch = 1;
    p2p_allreduce_kernel<<<nch, kP2PThreads, 0, static_cast<cudaStream_t>(stream)>>>(inout, n, reinterpret_cast<char* const*>(peer_bases), world, rank,
                                                                                     epoch, scale, n_pad, nch);
    count_launch();
    return int(cudaGetLastError());
}

}  // extern "C"
