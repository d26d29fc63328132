// FeatureExtractor, FVAE_PREC_BF16_TC: bf16 operands on tcgen05 tensor cores, fp32 accumulation in TMEM.
//
// Restates reference module.py:26-28 + the GRU input projection of :30 as ONE kernel per 128-row tile
// (a row = one (stock, time) pair of the panel):
//     x rows --LayerNorm(fp32)--> bf16 A tile --tcgen05.mma--> TMEM [128 x 160] (xn . W1^T)
//            --epilogue: +b1, LeakyReLU, bf16--> A tile 2 --tcgen05.mma--> TMEM [128 x pad16(3H)] (u . W_ih^T)
//            --epilogue: +b_ih--> gi[row][3H]
// so the panel is read once and neither xn nor u ever touch HBM.  Operand tiles use the chunk-major
// SWIZZLE_NONE layout of tc_sm100.cuh; weights are converted once per step into bf16 operand images.
//
// The GRU recurrence and the backward chain currently reuse the fp32 kernels of fe_f32.cu on the same
// workspace (gi / hall / dgh in fp32); they are being moved to tcgen05 tile by tile.
#include "fe.cuh"
#include "tc_sm100.cuh"

namespace fvae {

// shared with fe_f32.cu
struct FeF32Views { float *gi, *hall; };
FeF32Views fe_f32_views(const FeDims& d, void* ws);
int fe_f32_gru_forward(const FeDims& d, const FeW& w, void* ws, float* e, cudaStream_t st);

namespace {

using namespace tc;

constexpr int CP = 160;          // C padded to a multiple of 16 (K of both GEMMs, N of GEMM 1)
constexpr int KCH = CP / 8;      // 20 chunks of 8 features
constexpr int TM = 128;          // rows per tile = UMMA M
constexpr uint32_t A_BYTES = KCH * TM * 16;      // 40960
constexpr uint32_t W1_BYTES = KCH * CP * 16;     // 51200
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t ACC2_COL = 256;               // column of the second accumulator

inline int pad16(int v) { return (v + 15) & ~15; }

struct TcImages {      // bf16 operand images in global memory (chunk-major), rebuilt every step
    __nv_bfloat16* w1;     // [KCH][CP rows n][8]     B of GEMM 1: W1[n][k]
    __nv_bfloat16* wih;    // [KCH][N2 rows g][8]     B of GEMM 2: W_ih[g][k]
    int64_t bytes;
};

TcImages carve_images(const FeDims& d, void* base) {
    TcImages t;
    char* p = static_cast<char*>(base);
    t.w1 = reinterpret_cast<__nv_bfloat16*>(p);  p += W1_BYTES;
    t.wih = reinterpret_cast<__nv_bfloat16*>(p); p += size_t(KCH) * pad16(3 * d.H) * 16;
    t.bytes = ((p - static_cast<char*>(base)) + 255) / 256 * 256;
    return t;
}

// image[(k/8)][row][k%8] = W[row][k] (row-major fp32, ld = ldw), zero padded
__global__ void make_kmajor_image_kernel(const float* __restrict__ W, int rows, int cols, int ldw, int rows_pad, int kchunks,
                                         __nv_bfloat16* __restrict__ img) {
    const int total = kchunks * rows_pad * 8;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int e = idx & 7, row = (idx >> 3) % rows_pad, c = (idx >> 3) / rows_pad;
        const int k = c * 8 + e;
        const float v = (row < rows && k < cols) ? W[size_t(row) * ldw + k] : 0.f;
        img[idx] = __float2bfloat16(v);
    }
}

struct FrontArgs {
    const void* x; int x_bf16; int64_t seq_pitch, row_pitch; int contiguous;
    int T, C, H, N2; int64_t R;
    const float *ln_w, *ln_b, *b1, *bih;
    const __nv_bfloat16 *w1img, *wihimg;
    float* gi;     // [R][3H]
};

template <typename XT>
__device__ __forceinline__ float ld_stage(const unsigned char* row, int c) { return float(reinterpret_cast<const XT*>(row)[c]); }

// One CTA = 128 threads = 128 tile rows = 128 TMEM lanes.  Persistent over tiles.
template <typename XT>
__global__ void __launch_bounds__(TM, 1) fe_tc_front_fwd_kernel(FrontArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int C = a.C, H3 = 3 * a.H, N2 = a.N2;
    const uint32_t wih_bytes = uint32_t(KCH) * N2 * 16;
    // carve: weights | A1 | stage | A2 | vectors | barriers   (stage may spill into A2 for fp32 panels)
    unsigned char* sW1 = smem;
    unsigned char* sWih = sW1 + W1_BYTES;
    unsigned char* sA1 = sWih + wih_bytes;
    unsigned char* sStage = sA1 + A_BYTES;
    const uint32_t stage_bytes = (TM * C * uint32_t(sizeof(XT)) + 127u) & ~127u;
    unsigned char* sA2 = sStage + (sizeof(XT) == 2 ? stage_bytes : stage_bytes - A_BYTES);
    float* sVec = reinterpret_cast<float*>(sA2 + A_BYTES);     // gamma[CP] beta[CP] b1[CP] bih[N2]
    float* sGamma = sVec; float* sBeta = sVec + CP; float* sB1 = sVec + 2 * CP; float* sBih = sVec + 3 * CP;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sBih + N2);   // 2 mbarriers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

    // ---- one-time setup: weights -> smem, vectors, barriers, TMEM
    for (uint32_t i = tid; i < W1_BYTES / 16; i += TM) reinterpret_cast<uint4*>(sW1)[i] = reinterpret_cast<const uint4*>(a.w1img)[i];
    for (uint32_t i = tid; i < wih_bytes / 16; i += TM) reinterpret_cast<uint4*>(sWih)[i] = reinterpret_cast<const uint4*>(a.wihimg)[i];
    for (int i = tid; i < CP; i += TM) {
        sGamma[i] = i < C ? a.ln_w[i] : 0.f;
        sBeta[i] = i < C ? a.ln_b[i] : 0.f;
        sB1[i] = i < C ? a.b1[i] : 0.f;
    }
    for (int i = tid; i < N2; i += TM) sBih[i] = i < H3 ? a.bih[i] : 0.f;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc<TMEM_COLS>(tmem_slot);
    fence_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = uint32_t(warp) * 32u;
    const uint32_t idesc1 = make_idesc_bf16(TM, CP, false, false);
    const uint32_t idesc2 = make_idesc_bf16(TM, uint32_t(N2), false, false);
    const int64_t ntiles = (a.R + TM - 1) / TM;
    uint32_t phase = 0;

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, phase ^= 1) {
        const int64_t row0 = tile * TM;
        const int nrows = int((a.R - row0 < TM) ? a.R - row0 : TM);
        // ---- (a) raw rows -> stage
        if (a.contiguous && nrows == TM) {
            const uint4* src = reinterpret_cast<const uint4*>(static_cast<const XT*>(a.x) + row0 * C);
            const uint32_t n16 = TM * C * uint32_t(sizeof(XT)) / 16;
            for (uint32_t i = tid; i < n16; i += TM) reinterpret_cast<uint4*>(sStage)[i] = src[i];
        } else {
            for (int r = warp; r < TM; r += TM / 32) {
                XT* dst = reinterpret_cast<XT*>(sStage) + size_t(r) * C;
                if (r < nrows) {
                    const int64_t row = row0 + r;
                    const XT* src = static_cast<const XT*>(a.x) + (row / a.T) * a.seq_pitch + (row % a.T) * a.row_pitch;
                    for (int c = tid & 31; c < C; c += 32) dst[c] = src[c];
                } else {
                    for (int c = tid & 31; c < C; c += 32) dst[c] = XT(0.f);
                }
            }
        }
        __syncthreads();
        // ---- (b) LayerNorm of my row (fp32 statistics over exactly C features) -> bf16 A1
        {
            const unsigned char* rowp = sStage + size_t(tid) * C * sizeof(XT);
            float sum = 0.f;
            for (int c = 0; c < C; ++c) sum += ld_stage<XT>(rowp, c);
            const float mean = sum / float(C);
            float sq = 0.f;
            for (int c = 0; c < C; ++c) { const float dlt = ld_stage<XT>(rowp, c) - mean; sq = fmaf(dlt, dlt, sq); }
            const float rstd = rsqrtf(sq / float(C) + kLnEps);
#pragma unroll 1
            for (int ch = 0; ch < KCH; ++ch) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int c = ch * 8 + e;
                    v[e] = (c < C) ? fmaf((ld_stage<XT>(rowp, c) - mean) * rstd, sGamma[c], sBeta[c]) : 0.f;
                }
                uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                *reinterpret_cast<uint4*>(sA1 + tile_off(TM, tid, ch)) = pk;
            }
        }
        fence_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        // ---- (c) GEMM 1: [128 x 160] = A1 . W1^T
        if (tid == 0) {
            tc_fence_after_sync();
            const uint32_t a0 = smem_u32(sA1), b0 = smem_u32(sW1);
#pragma unroll
            for (int ks = 0; ks < CP / 16; ++ks) {
                const uint64_t ad = make_smem_desc(a0 + ks * 2 * (TM * 16), TM * 16, 128);
                const uint64_t bd = make_smem_desc(b0 + ks * 2 * (CP * 16), CP * 16, 128);
                mma_bf16_ss(tmem, ad, bd, idesc1, ks > 0);
            }
            mma_commit(&bars[0]);
        }
        mbar_wait(&bars[0], phase);
        tc_fence_after_sync();
        // ---- (d) epilogue 1: +b1, LeakyReLU, bf16 -> A2
#pragma unroll 1
        for (int j = 0; j < CP / 16; ++j) {
            float v[16];
            tmem_ld16(tmem_addr(tmem, lane_base, j * 16), v);
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = lrelu(v[e] + sB1[j * 16 + e]);
            uint4 p0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            uint4 p1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
            *reinterpret_cast<uint4*>(sA2 + tile_off(TM, tid, 2 * j)) = p0;
            *reinterpret_cast<uint4*>(sA2 + tile_off(TM, tid, 2 * j + 1)) = p1;
        }
        fence_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        // ---- (e) GEMM 2: [128 x N2] = A2 . W_ih^T
        if (tid == 0) {
            tc_fence_after_sync();
            const uint32_t a0 = smem_u32(sA2), b0 = smem_u32(sWih);
#pragma unroll
            for (int ks = 0; ks < CP / 16; ++ks) {
                const uint64_t ad = make_smem_desc(a0 + ks * 2 * (TM * 16), TM * 16, 128);
                const uint64_t bd = make_smem_desc(b0 + ks * 2 * (uint32_t(N2) * 16), uint32_t(N2) * 16, 128);
                mma_bf16_ss(tmem + ACC2_COL, ad, bd, idesc2, ks > 0);
            }
            mma_commit(&bars[1]);
        }
        mbar_wait(&bars[1], phase);
        tc_fence_after_sync();
        // ---- (f) epilogue 2: +b_ih -> gi
        {
            float* out = a.gi + (row0 + tid) * H3;
#pragma unroll 1
            for (int j = 0; j < N2 / 16; ++j) {
                float v[16];
                tmem_ld16(tmem_addr(tmem, lane_base, ACC2_COL + j * 16), v);
                if (tid < nrows) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int n = j * 16 + e;
                        if (n < H3) out[n] = v[e] + sBih[n];
                    }
                }
            }
        }
        tc_fence_before_sync();
        __syncthreads();     // stage / A1 / TMEM may be overwritten by the next tile
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem);
}

size_t front_smem_bytes(int C, int N2, size_t esize) {
    const size_t stage = (size_t(TM) * C * esize + 127) & ~size_t(127);
    size_t b = W1_BYTES + size_t(KCH) * N2 * 16 + A_BYTES;
    b += (esize == 2) ? stage + A_BYTES : stage;       // fp32 staging overlaps A2
    b += (3 * CP + N2) * sizeof(float) + 2 * sizeof(uint64_t) + 16;
    return b;
}

}  // namespace

int fe_tc_supported(const FeDims& d) {
    if (d.C > CP || d.C < 16 || d.H > kMaxH) return FVAE_ERR_UNSUPPORTED;
    return 0;
}

int64_t fe_tc_workspace_bytes(const FeDims& d) {
    const int64_t f32 = (fe_f32_workspace_bytes(d) + 255) / 256 * 256;
    return f32 + carve_images(d, nullptr).bytes;
}

int fe_tc_forward(const FeDims& d, const fvae_panel& x, const FeW& w, float* e, void* ws, cudaStream_t st) {
    const int64_t f32 = (fe_f32_workspace_bytes(d) + 255) / 256 * 256;
    TcImages img = carve_images(d, static_cast<char*>(ws) + f32);
    FeF32Views v = fe_f32_views(d, ws);
    const int N2 = pad16(3 * d.H);
    make_kmajor_image_kernel<<<32, 256, 0, st>>>(w.W1, d.C, d.C, d.C, CP, KCH, img.w1); count_launch();
    make_kmajor_image_kernel<<<32, 256, 0, st>>>(w.Wih, 3 * d.H, d.C, d.C, N2, KCH, img.wih); count_launch();
    FrontArgs a;
    a.x = x.data; a.x_bf16 = (x.dtype == FVAE_BF16); a.seq_pitch = x.seq_pitch; a.row_pitch = x.row_pitch;
    a.contiguous = (x.row_pitch == d.C && x.seq_pitch == int64_t(d.T) * d.C && (reinterpret_cast<uintptr_t>(x.data) % 16 == 0));
    a.T = d.T; a.C = d.C; a.H = d.H; a.N2 = N2; a.R = int64_t(d.S) * d.T;
    a.ln_w = w.ln_w; a.ln_b = w.ln_b; a.b1 = w.b1; a.bih = w.bih;
    a.w1img = img.w1; a.wihimg = img.wih; a.gi = v.gi;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ntiles = (a.R + TM - 1) / TM;
    const int grid = int(ntiles < nsm ? ntiles : nsm);
    cudaError_t ce;
    if (a.x_bf16) {
        const size_t smem = front_smem_bytes(d.C, N2, 2);
        if (smem > 227 * 1024) return FVAE_ERR_LIMIT;
        if ((ce = cudaFuncSetAttribute(fe_tc_front_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))) != cudaSuccess) return int(ce);
        fe_tc_front_fwd_kernel<__nv_bfloat16><<<grid, TM, smem, st>>>(a); count_launch();
    } else {
        const size_t smem = front_smem_bytes(d.C, N2, 4);
        if (smem > 227 * 1024) return FVAE_ERR_LIMIT;
        if ((ce = cudaFuncSetAttribute(fe_tc_front_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))) != cudaSuccess) return int(ce);
        fe_tc_front_fwd_kernel<float><<<grid, TM, smem, st>>>(a); count_launch();
    }
    if ((ce = cudaGetLastError()) != cudaSuccess) return int(ce);
    return fe_f32_gru_forward(d, w, ws, e, st);
}

int fe_tc_backward(const FeDims& d, const fvae_panel& x, const FeW& w, const FeG& g, const float* dE, void* ws, cudaStream_t st) {
    // TODO(tcgen05): BPTT and the front backward still run the fp32 kernels on the shared workspace.
    return fe_f32_backward(d, x, w, g, dE, ws, st);
}

}  // namespace fvae
