// Generator front end on the tensor cores (16-bit modes, ndomain 16):
//   Flatten + Concatenate([latent(100), cond(256)])  gan_train_cwgangp_pixelnorm.py:319-323
//   Dense(3072) + LeakyReLU(0.2) + Reshape((3,2,2,256))                        :326-328
// as ONE kernel: x0 is never materialised, the output is written as the 16-bit channels-last tensor the first
// upsampled conv reads with TMA (flat index ((t*2+h)*2+w)*256+c == Dense output index).
//
// GEMM shape: D[128 samples, 256 outputs] = X0[128, 384] * W^T[384, 256]  (K = 356 zero-padded to 6 chunks of 64).
// A CTA owns one block of 128 samples and a group of 3 output blocks: the A operand (concat, converted to 16 bit,
// 128B-swizzled K-major) is built once in shared memory by all threads; the 32 KB weight tiles stream through a
// 3-deep bulk-copy ring; N = 256 MMAs are math-bound even from a single issuing thread (tools/umma_rate_probe.cu).
// Warp roles (192 threads): 0 = weight producer, 1 = TMEM alloc + MMA issuer, 2..5 = epilogue.
#include "rdg_common.cuh"
#include "gen_tc.h"
#include "tc_ptx.cuh"

using namespace rdg_tc;

namespace {

constexpr int kThreads = 192;
constexpr int kKin = 356, kChunks = 6;            // K padded to 384
constexpr int kNOut = 3072, kNBlk = 256, kNBlocks = kNOut / kNBlk, kNbPerCta = 3, kGroups = kNBlocks / kNbPerCta;
constexpr int kATile = 128 * 128;                 // one 64-wide K chunk of the sample block
constexpr int kBTile = kNBlk * 128;               // 32 KB: 256 output rows x 64 k
constexpr int kBStages = 3;
constexpr int kSmem = 1024 + kChunks * kATile + kBStages * kBTile + 256;
static_assert(kSmem <= 227 * 1024, "shared memory overflow");

template <typename HT>
__global__ void __launch_bounds__(kThreads, 1)
tc_dense_lrelu_kernel(const float* __restrict__ latent, const float* __restrict__ cond, int spc, int b_off,
                      const HT* __restrict__ wpack, const float* __restrict__ bias, HT* __restrict__ out, int B) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_buf = smem;
    uint8_t* b_buf = a_buf + kChunks * kATile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + kBStages * kBTile);
    uint64_t* b_full = bars;
    uint64_t* b_empty = b_full + kBStages;
    uint64_t* acc_full = b_empty + kBStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    __shared__ float s_bias[kNbPerCta * kNBlk];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mb = blockIdx.x / kGroups, grp = blockIdx.x % kGroups;
    const int row0 = mb * 128;

    for (int i = threadIdx.x; i < kNbPerCta * kNBlk; i += kThreads) s_bias[i] = bias[grp * kNbPerCta * kNBlk + i];
    // A operand: x0[row][k] = k < 100 ? latent[row][k] : cond[(b_off + row) / spc][k - 100]; rows >= B and k >= 356 are zero.
    // Chunk c holds k in [64c, 64c+64); 16-byte group j of row r lives at group (j ^ (r & 7)) (SWIZZLE_128B, K-major).
    // Every 8-element group splits into two float4 halves, each entirely latent, cond or zero padding (100 and 356 are
    // multiples of 4); consecutive threads take consecutive groups of one row, so the global reads coalesce.
    for (int g = threadIdx.x; g < 128 * kChunks * 8; g += kThreads) {
        const int r = g / (kChunks * 8), jg = g % (kChunks * 8);   // jg = chunk*8 + j
        const int c = jg >> 3, j = jg & 7;
        const int row = row0 + r;
        float4 h[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int k = jg * 8 + q * 4;
            h[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < B) {
                if (k < RDG_LATENT) h[q] = *reinterpret_cast<const float4*>(latent + (size_t)row * RDG_LATENT + k);
                else if (k < kKin) h[q] = *reinterpret_cast<const float4*>(cond + (size_t)((b_off + row) / spc) * (kKin - RDG_LATENT) + (k - RDG_LATENT));
            }
        }
        *reinterpret_cast<uint4*>(a_buf + c * kATile + r * 128 + ((j ^ (r & 7)) << 4)) =
            make_uint4(HalfOps<HT>::pack(h[0].x, h[0].y), HalfOps<HT>::pack(h[0].z, h[0].w),
                       HalfOps<HT>::pack(h[1].x, h[1].y), HalfOps<HT>::pack(h[1].z, h[1].w));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int i = 0; i < kBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // weight producer: tile (output block nb, chunk c) is one contiguous 32 KB image
        uint32_t bi = 0;
        for (int i = 0; i < kNbPerCta; ++i)
            for (int c = 0; c < kChunks; ++c, ++bi) {
                const uint32_t s = bi % kBStages, ph = (bi / kBStages) & 1;
                mbar_wait(&b_empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&b_full[s], kBTile);
                    bulk_load_1d(b_buf + s * kBTile,
                                 reinterpret_cast<const uint8_t*>(wpack) + ((size_t)(grp * kNbPerCta + i) * kChunks + c) * kBTile, kBTile, &b_full[s]);
                }
                __syncwarp();
            }
    } else if (warp == 1) {
        constexpr uint32_t kF = HalfOps<HT>::kFmt;
        constexpr uint32_t idesc = (1u << 4) | (kF << 7) | (kF << 10) | ((uint32_t)(kNBlk >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint32_t bi = 0;
        for (int i = 0; i < kNbPerCta; ++i) {
            const uint32_t as = i & 1, aph = (i >> 1) & 1;
            mbar_wait(&acc_empty[as], aph ^ 1);
            tc_fence_after();
            for (int c = 0; c < kChunks; ++c, ++bi) {
                const uint32_t s = bi % kBStages, ph = (bi / kBStages) & 1;
                mbar_wait(&b_full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t ad = make_sdesc(smem_u32(a_buf + c * kATile)), bd = make_sdesc(smem_u32(b_buf + s * kBTile));
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base + as * kNBlk, ad + 2 * k, bd + 2 * k, idesc, (c | k) ? 1u : 0u);
                    tc_commit(&b_empty[s]);
                    if (c == kChunks - 1) tc_commit(&acc_full[as]);
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;                    // warps 2..5 -> TMEM lane quarters 2,3,0,1
        const int r = q * 32 + lane;
        const int row = row0 + r;
        for (int i = 0; i < kNbPerCta; ++i) {
            const uint32_t as = i & 1, aph = (i >> 1) & 1;
            mbar_wait(&acc_full[as], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * kNBlk;
            HT* orow = out + (size_t)row * kNOut + (size_t)(grp * kNbPerCta + i) * kNBlk;
#pragma unroll 1
            for (int c0 = 0; c0 < kNBlk; c0 += 32) {
                uint32_t v[32];
                tc_ld32(taddr + c0, v);
                if (row < B) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float x0 = __uint_as_float(v[2 * j]) + s_bias[i * kNBlk + c0 + 2 * j];
                        float x1 = __uint_as_float(v[2 * j + 1]) + s_bias[i * kNBlk + c0 + 2 * j + 1];
                        x0 = x0 > 0.f ? x0 : 0.2f * x0;
                        x1 = x1 > 0.f ? x1 : 0.2f * x1;
                        pk[j] = HalfOps<HT>::pack(x0, x1);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(orow + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// Dense kernel (356, 3072) f32, Keras (in, out) order -> [12 output blocks][6 chunks][256 rows x 64 k] 16-bit, swizzled, k >= 356 zero
template <typename HT>
__global__ void pack_dense_kernel(const float* __restrict__ w, HT* __restrict__ dst) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= kNBlocks * kChunks * kNBlk * 64) return;
    const int pos = idx & 63, n = (idx >> 6) % kNBlk, tile = idx / (64 * kNBlk);
    const int c = tile % kChunks, nb = tile / kChunks;
    const int k = c * 64 + (((pos >> 3) ^ (n & 7)) << 3) + (pos & 7);
    dst[idx] = HalfOps<HT>::from_float(k < kKin ? w[(size_t)k * kNOut + nb * kNBlk + n] : 0.f);
}

template <typename HT>
int launch_dense(const float* latent, const float* cond, int spc, int b_off, const void* wpack, const float* bias, void* out,
                 int B, cudaStream_t st) {
    auto kern = tc_dense_lrelu_kernel<HT>;
    static bool attr_set = false;
    if (!attr_set) {
        RDG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_set = true;
    }
    kern<<<ceil_div(B, 128) * kGroups, kThreads, kSmem, st>>>(latent, cond, spc, b_off, reinterpret_cast<const HT*>(wpack), bias,
                                                             reinterpret_cast<HT*>(out), B);
    RDG_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int tc_dense_lrelu(int half_kind, const float* latent, const float* cond, int spc, int b_off, const void* wpack,
                   const float* bias, void* out, int B, cudaStream_t st) {
    if (B <= 0) return 0;
    if (half_kind == RDG_HALF_BF16) return launch_dense<__nv_bfloat16>(latent, cond, spc, b_off, wpack, bias, out, B, st);
    return launch_dense<__half>(latent, cond, spc, b_off, wpack, bias, out, B, st);
}

size_t tc_dense_pack_bytes() { return (size_t)kNBlocks * kChunks * kBTile; }

int pack_dense_weights(int half_kind, const float* w, void* dst, cudaStream_t st) {
    const int n = kNBlocks * kChunks * kNBlk * 64;
    if (half_kind == RDG_HALF_BF16) pack_dense_kernel<__nv_bfloat16><<<ceil_div(n, 256), 256, 0, st>>>(w, (__nv_bfloat16*)dst);
    else pack_dense_kernel<__half><<<ceil_div(n, 256), 256, 0, st>>>(w, (__half*)dst);
    RDG_LAUNCH_CHECK();
    return 0;
}
