// Generator front end on the tensor cores for the shapes the small fused kernel (gen_dense_tc.cu) does not cover: the
// large-domain variant (alternative_domains/gan_train_cwgangp_pixelnorm_largedomain.py:323-335: Dense 4196 -> 49152 at
// ndomain 64) and the additional-input variants (ncond > 1):
//   Flatten + Concatenate([latent, cond]) -> Dense(256*(nd/8)^2*3) + LeakyReLU(0.2) -> Reshape((3, nd/8, nd/8, 256)).
// A plain GEMM  D[B, N] = X0[B, K] * W[K, N]: X0 is assembled and converted to 16 bit by a small kernel ([B, Kp], Kp = K
// rounded up to 64, zero padded), W is packed once per weight update as its 16-bit transpose [N, Kp]; both are K-major, so
// 2-D TMA boxes (64 k x 128 rows / 64 k x 256 rows, SWIZZLE_128B) land MMA-ready tiles.  One CTA = one 128 x 256 output tile,
// 4-stage TMA ring over the Kp / 64 chunks, tcgen05.mma M=128 N=256 K=16 accumulating in TMEM, epilogue = bias + LeakyReLU +
// 16-bit store in the channels-last layout the first upsampled conv reads (flat Dense index == ((t*s+h)*s+w)*256+c).
// Weight-bound at small batch (412 MB of 16-bit weights per pass at nd = 64), math-bound from ~300 samples per pass.
// Warp roles (192 threads): 0 = TMA producer, 1 = TMEM alloc + MMA issuer, 2..5 = epilogue.
#include "rdg_common.cuh"
#include "gen_tc.h"
#include "tc_ptx.cuh"
#include <cuda.h>

using namespace rdg_tc;

namespace {

constexpr int kThreads = 192;
constexpr int kBM = 128, kBN = 256;
constexpr int kATile = kBM * 128, kBTile = kBN * 128;      // 16 KB + 32 KB per 64-wide K chunk
constexpr int kStages = 4;
constexpr int kSmem = 1024 + kStages * (kATile + kBTile) + 256;
static_assert(kSmem <= 227 * 1024, "shared memory overflow");

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

template <typename HT>
__global__ void __launch_bounds__(kThreads, 1)
tc_dense_big_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const float* __restrict__ bias, HT* __restrict__ out, int B, int N, int nk) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_buf = smem;
    uint8_t* b_buf = a_buf + kStages * kATile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + kStages * kBTile);
    uint64_t* full = bars;
    uint64_t* empty = full + kStages;
    uint64_t* acc_full = empty + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    __shared__ float s_bias[kBN];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * kBN, m0 = blockIdx.y * kBM;        // N tiles fastest: CTAs of a wave share the X rows in L2
    for (int i = threadIdx.x; i < kBN; i += kThreads) s_bias[i] = bias[n0 + i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        for (int kc = 0; kc < nk; ++kc) {
            const uint32_t s = kc % kStages, ph = (kc / kStages) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&full[s], kATile + kBTile);
                tma_load_2d(a_buf + s * kATile, &tmap_x, &full[s], kc * 64, m0);      // rows >= B: zero fill
                tma_load_2d(b_buf + s * kBTile, &tmap_w, &full[s], kc * 64, n0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t kF = HalfOps<HT>::kFmt;
        constexpr uint32_t idesc = (1u << 4) | (kF << 7) | (kF << 10) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
        for (int kc = 0; kc < nk; ++kc) {
            const uint32_t s = kc % kStages, ph = (kc / kStages) & 1;
            mbar_wait(&full[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t ad = make_sdesc(smem_u32(a_buf + s * kATile)), bd = make_sdesc(smem_u32(b_buf + s * kBTile));
#pragma unroll
                for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, (kc | k) ? 1u : 0u);
                tc_commit(&empty[s]);
                if (kc == nk - 1) tc_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3;                      // TMEM lane quarter this warp may touch
        const int row = m0 + q * 32 + lane;
        mbar_wait(acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < kBN; c0 += 32) {
            uint32_t v[32];
            tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
            if (row < B) {
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float x0 = __uint_as_float(v[2 * j]) + s_bias[c0 + 2 * j], x1 = __uint_as_float(v[2 * j + 1]) + s_bias[c0 + 2 * j + 1];
                    x0 = x0 > 0.f ? x0 : 0.2f * x0; x1 = x1 > 0.f ? x1 : 0.2f * x1;
                    pk[j] = HalfOps<HT>::pack(x0, x1);
                }
                uint4* dst = reinterpret_cast<uint4*>(out + (size_t)row * N + n0 + c0);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// x16[b][k] = k < 100 ? latent[b][k] : k < K ? cond[(b_off + b) / spc][k - 100] : 0
template <typename HT>
__global__ void dense_big_input_kernel(const float* __restrict__ latent, const float* __restrict__ cond, int spc, int b_off,
                                       HT* __restrict__ x16, int B, int K, int Kp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * Kp) return;
    const int b = (int)(i / Kp), k = (int)(i % Kp);
    float v = 0.f;
    if (k < RDG_LATENT) v = latent[(size_t)b * RDG_LATENT + k];
    else if (k < K) v = cond[(size_t)((b_off + b) / spc) * (K - RDG_LATENT) + (k - RDG_LATENT)];
    x16[i] = HalfOps<HT>::from_float(v);
}

// wT[n][k] = W[k][n] (Keras Dense kernel is [K, N]), zero for K <= k < Kp; 32 x 32 tiles through shared memory
template <typename HT>
__global__ void dense_big_pack_kernel(const float* __restrict__ w, HT* __restrict__ wT, int K, int N, int Kp) {
    __shared__ float t[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int k = k0 + r, n = n0 + threadIdx.x;
        t[r][threadIdx.x] = (k < K && n < N) ? w[(size_t)k * N + n] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
        const int n = n0 + r, k = k0 + threadIdx.x;
        if (n < N && k < Kp) wT[(size_t)n * Kp + k] = HalfOps<HT>::from_float(t[threadIdx.x][r]);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

template <typename HT>
int launch_dense_big(const float* latent, const float* cond, int spc, int b_off, const void* wT, const float* bias, void* x16,
                     void* out, int B, int K, int N, cudaStream_t st) {
    const int Kp = (K + 63) / 64 * 64;
    if (N % kBN) { rdg_set_error("tc dense: N must be a multiple of %d", kBN); return RDG_TC_E_SHAPE; }
    EncodeTiledFn enc = encode_fn();
    if (!enc) { rdg_set_error("cuTensorMapEncodeTiled entry point not available"); return RDG_TC_E_DRIVER; }
    dense_big_input_kernel<HT><<<ceil_div((long long)B * Kp, 256), 256, 0, st>>>(latent, cond, spc, b_off, (HT*)x16, B, K, Kp);
    RDG_LAUNCH_CHECK();
    const CUtensorMapDataType dt = HalfOps<HT>::kFmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap tx, tw;
    cuuint32_t es[2] = {1, 1};
    {
        cuuint64_t gd[2] = {(cuuint64_t)Kp, (cuuint64_t)B}; cuuint64_t gs[1] = {(cuuint64_t)Kp * 2}; cuuint32_t box[2] = {64, kBM};
        CUresult r = enc(&tx, dt, 2, x16, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { rdg_set_error("cuTensorMapEncodeTiled (dense x) failed: %d", (int)r); return RDG_TC_E_DRIVER; }
    }
    {
        cuuint64_t gd[2] = {(cuuint64_t)Kp, (cuuint64_t)N}; cuuint64_t gs[1] = {(cuuint64_t)Kp * 2}; cuuint32_t box[2] = {64, kBN};
        CUresult r = enc(&tw, dt, 2, const_cast<void*>(wT), gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { rdg_set_error("cuTensorMapEncodeTiled (dense w) failed: %d", (int)r); return RDG_TC_E_DRIVER; }
    }
    auto kern = tc_dense_big_kernel<HT>;
    static bool attr_set = false;
    if (!attr_set) { RDG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem)); attr_set = true; }
    kern<<<dim3(N / kBN, ceil_div(B, kBM)), kThreads, kSmem, st>>>(tx, tw, bias, (HT*)out, B, N, Kp / 64);
    RDG_LAUNCH_CHECK();
    return 0;
}

}  // namespace

size_t tc_dense_big_pack_bytes(int K, int N) { return (size_t)N * ((K + 63) / 64 * 64) * 2; }
size_t tc_dense_big_input_bytes(int B, int K) { return (size_t)B * ((K + 63) / 64 * 64) * 2; }

int pack_dense_big_weights(int half_kind, const float* w, void* dst, int K, int N, cudaStream_t st) {
    const int Kp = (K + 63) / 64 * 64;
    dim3 grid(ceil_div(Kp, 32), ceil_div(N, 32)), block(32, 8);
    if (half_kind == RDG_HALF_BF16) dense_big_pack_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(w, (__nv_bfloat16*)dst, K, N, Kp);
    else dense_big_pack_kernel<__half><<<grid, block, 0, st>>>(w, (__half*)dst, K, N, Kp);
    RDG_LAUNCH_CHECK();
    return 0;
}

int tc_dense_big_lrelu(int half_kind, const float* latent, const float* cond, int spc, int b_off, const void* wT, const float* bias,
                       void* x16_scratch, void* out, int B, int K, int N, cudaStream_t st) {
    if (B <= 0) return 0;
    if (half_kind == RDG_HALF_BF16)
        return launch_dense_big<__nv_bfloat16>(latent, cond, spc, b_off, wT, bias, x16_scratch, out, B, K, N, st);
    return launch_dense_big<__half>(latent, cond, spc, b_off, wT, bias, x16_scratch, out, B, K, N, st);
}
