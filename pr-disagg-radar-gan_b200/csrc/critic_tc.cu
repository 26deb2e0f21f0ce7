// Critic forward on the tensor cores (16-bit scoring mode): create_discriminator, gan_train_cwgangp_pixelnorm.py:272-309.
//   D1  Conv3D(64, 3, strides=2, 'valid') on the 2-channel input (:286-287): K = 27 * (1 + ncond) is tiny -> CUDA cores,
//       fused with the condition tiling / concat (:275-282) and LeakyReLU, 16-bit channels-last output;
//   D2..D4  Conv3D(128 / 256 / 256, 3, strides=2, 'same') + LeakyReLU (:291-301) as implicit GEMMs on tcgen05:
//       D[128 output positions, Cout] += A[128, 64] * W[tap][64-channel chunk][64, Cout] over 27 taps x Cin/64 chunks.
//       A = the input window of the tap, fetched by ONE 5-D TMA box load with element strides (1, 2, 2, 2, 1): box
//       (64 ch, Wo, Hb, 1, Bt) output positions -> 128 rows of 128 B, SWIZZLE_128B; coordinates start at 2*o + k - pad, and
//       out-of-range elements are zero filled, which is exactly TF 'same' padding including its asymmetric (0,1) case.
//       B = pre-swizzled weight tiles (1-D bulk copies); accumulators in TMEM, double buffered; epilogue bias + LeakyReLU +
//       16-bit store.  Dropout is identity in scoring (inference) mode;
//   D5  Flatten + Dense(1) (:303-304): one warp per sample on CUDA cores.
// Warp roles (224 threads): 0 = A producer, 1 = TMEM alloc + MMA issuer, 2 = weight producer, 3..6 = epilogue.
#include "rdg_common.cuh"
#include "gen_tc.h"
#include "tc_ptx.cuh"
#include <cuda.h>

using namespace rdg_tc;

namespace {

constexpr int kThreads = 224;
constexpr int kATile = 128 * 128;            // 128 rows x 64 ch x 2 B
constexpr int kAStages = 4;

template <int COUT> struct CriticCfg {
    static constexpr int kBTile = COUT * 128;
    static constexpr int kBStages = (200 * 1024 - kAStages * kATile) / kBTile > 6 ? 6 : (200 * 1024 - kAStages * kATile) / kBTile;
    static constexpr int kSmem = 1024 + kAStages * kATile + kBStages * kBTile + COUT * 4 + 512;
    static_assert(kBStages >= 2 && kSmem <= 227 * 1024, "shared memory");
};

struct CriticConvArgs {
    int B, To, Ho, Wo, Cin;       // output grid, input channels
    int Hb, Bt;                   // tile = Bt samples x Hb rows x Wo columns of one output hour plane = 128 rows
    int pt, ph, pw;               // zero padding before (TF 'same')
    int n_tiles;
    const void* wpack;            // [27 taps][Cin/64][Cout rows x 64 k] 16-bit, rows 128B-swizzled
    const float* bias;
    void* out;                    // [B,To,Ho,Wo,Cout] 16-bit
};

template <typename HT, int COUT>
__global__ void __launch_bounds__(kThreads, 1)
tc_critic_conv_kernel(const __grid_constant__ CUtensorMap tmap, CriticConvArgs args) {
    using Cfg = CriticCfg<COUT>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_buf = smem;
    uint8_t* b_buf = a_buf + kAStages * kATile;
    float* s_bias = reinterpret_cast<float*>(b_buf + Cfg::kBStages * Cfg::kBTile);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + COUT);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + kAStages;
    uint64_t* b_full = a_empty + kAStages;
    uint64_t* b_empty = b_full + Cfg::kBStages;
    uint64_t* acc_full = b_empty + Cfg::kBStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunk = args.Cin / 64, n_hblk = args.Ho / args.Hb;
    const int nsteps = 27 * nchunk;

    for (int i = threadIdx.x; i < COUT; i += kThreads) s_bias[i] = args.bias[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < kAStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < Cfg::kBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
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
        // ================= A producer: one strided box per (tap, 64-channel chunk) =================
        uint32_t ai = 0;
        for (int tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x) {
            const int hblk = tile % n_hblk, t = (tile / n_hblk) % args.To, bblk = tile / (n_hblk * args.To);
            const int h0 = hblk * args.Hb, b0 = bblk * args.Bt;
            for (int tap = 0; tap < 27; ++tap) {
                const int kt = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
                for (int c = 0; c < nchunk; ++c, ++ai) {
                    const uint32_t s = ai % kAStages, ph = (ai / kAStages) & 1;
                    mbar_wait(&a_empty[s], ph ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&a_full[s], kATile);
                        tma_load_5d(a_buf + s * kATile, &tmap, &a_full[s], c * 64, kw - args.pw, 2 * h0 + kh - args.ph,
                                    2 * t + kt - args.pt, b0);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 2) {
        // ================= weight producer =================
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(args.wpack);
        uint32_t bi = 0;
        for (int tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x)
            for (int st = 0; st < nsteps; ++st, ++bi) {
                const uint32_t s = bi % Cfg::kBStages, ph = (bi / Cfg::kBStages) & 1;
                mbar_wait(&b_empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&b_full[s], Cfg::kBTile);
                    bulk_load_1d(b_buf + s * Cfg::kBTile, wsrc + (size_t)st * Cfg::kBTile, Cfg::kBTile, &b_full[s]);
                }
                __syncwarp();
            }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        constexpr uint32_t kF = HalfOps<HT>::kFmt;
        constexpr uint32_t idesc = (1u << 4) | (kF << 7) | (kF << 10) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        uint32_t ai = 0, bi = 0, acc_it = 0;
        for (int tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x, ++acc_it) {
            const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
            mbar_wait(&acc_empty[as], aph ^ 1);
            tc_fence_after();
            const uint32_t d_acc = tmem_base + as * COUT;
            for (int st = 0; st < nsteps; ++st, ++ai, ++bi) {
                const uint32_t sa = ai % kAStages, pa = (ai / kAStages) & 1;
                const uint32_t sb = bi % Cfg::kBStages, pb = (bi / Cfg::kBStages) & 1;
                mbar_wait(&a_full[sa], pa);
                mbar_wait(&b_full[sb], pb);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t ad = make_sdesc(smem_u32(a_buf + sa * kATile)), bd = make_sdesc(smem_u32(b_buf + sb * Cfg::kBTile));
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc_mma_f16(d_acc, ad + 2 * k, bd + 2 * k, idesc, (st | k) ? 1u : 0u);
                    tc_commit(&a_empty[sa]);
                    tc_commit(&b_empty[sb]);
                    if (st == nsteps - 1) tc_commit(&acc_full[as]);
                }
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue: bias + LeakyReLU(0.2) -> 16-bit channels-last =================
        const int q = warp & 3;
        const int r = q * 32 + lane;                           // accumulator row = ((b_local * Hb) + h_local) * Wo + w
        const int bl = r / (args.Hb * args.Wo), hl = (r / args.Wo) % args.Hb, w = r % args.Wo;
        HT* out = reinterpret_cast<HT*>(args.out);
        uint32_t acc_it = 0;
        for (int tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x, ++acc_it) {
            const int hblk = tile % n_hblk, t = (tile / n_hblk) % args.To, bblk = tile / (n_hblk * args.To);
            const int b = bblk * args.Bt + bl, h = hblk * args.Hb + hl;
            const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
            mbar_wait(&acc_full[as], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * COUT;
            const size_t o_pos = (((size_t)b * args.To + t) * args.Ho + h) * args.Wo + w;
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                uint32_t v[32];
                tc_ld32(taddr + c0, v);
                if (b < args.B) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float x0 = __uint_as_float(v[2 * j]) + s_bias[c0 + 2 * j], x1 = __uint_as_float(v[2 * j + 1]) + s_bias[c0 + 2 * j + 1];
                        x0 = x0 > 0.f ? x0 : 0.2f * x0; x1 = x1 > 0.f ? x1 : 0.2f * x1;
                        pk[j] = HalfOps<HT>::pack(x0, x1);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(out + o_pos * COUT + c0);
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

// D1: cond tiling + concat + Conv3D(64, 3, s=2, 'valid') + LeakyReLU on CUDA cores; one thread per (output position, 8 channels)
template <typename HT>
__global__ void __launch_bounds__(256) critic_first_conv_kernel(const float* __restrict__ sample, const float* __restrict__ cond,
                                                                const float* __restrict__ w, const float* __restrict__ bias,
                                                                HT* __restrict__ out, int B, int nd, int ncond, int To, int Ho, int Wo) {
    extern __shared__ float ws[];                               // [27][Cin][64]
    const int Cin = 1 + ncond;
    for (int i = threadIdx.x; i < 27 * Cin * 64; i += blockDim.x) ws[i] = w[i];
    __syncthreads();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * To * Ho * Wo * 8;
    if (idx >= total) return;
    const int cg = (int)(idx & 7);
    long long r = idx >> 3;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho); r /= Ho;
    const int to = (int)(r % To); const int b = (int)(r / To);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias[cg * 8 + j];
    for (int kt = 0; kt < 3; ++kt)
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                const int t = 2 * to + kt, h = 2 * ho + kh, x = 2 * wo + kw;          // 'valid': always inside
                const int tap = (kt * 3 + kh) * 3 + kw;
                const float s0 = sample[(((size_t)b * RDG_NHOURS + t) * nd + h) * nd + x];
                const float* wt = ws + (size_t)tap * Cin * 64 + cg * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(s0, wt[j], acc[j]);
                for (int c = 0; c < ncond; ++c) {                                    // condition channels, tiled over the hours
                    const float cv = cond[(((size_t)b * nd + h) * nd + x) * ncond + c];
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(cv, wt[(c + 1) * 64 + j], acc[j]);
                }
            }
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float x0 = acc[2 * j], x1 = acc[2 * j + 1];
        x0 = x0 > 0.f ? x0 : 0.2f * x0; x1 = x1 > 0.f ? x1 : 0.2f * x1;
        pk[j] = HalfOps<HT>::pack(x0, x1);
    }
    *reinterpret_cast<uint4*>(out + (idx >> 3) * 64 + cg * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// D5: score[b] = dot(flatten(h4[b]), w5) + b5; one warp per sample
template <typename HT>
__global__ void __launch_bounds__(256) critic_dense_kernel(const HT* __restrict__ h4, const float* __restrict__ w5, const float* __restrict__ b5,
                                                           float* __restrict__ score, int B, int K) {
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    float s = 0.f;
    for (int k = lane * 2; k < K; k += 64) {
        const float2 v = HalfOps<HT>::unpack(*reinterpret_cast<const uint32_t*>(h4 + (size_t)b * K + k));
        s = fmaf(v.x, w5[k], fmaf(v.y, w5[k + 1], s));
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) score[b] = s + b5[0];
}

// wpack[tap][chunk][n][k'] = W[tap][chunk*64 + k][n], rows 128B-swizzled (16-byte group j of row n at j ^ (n & 7))
template <typename HT>
__global__ void critic_pack_kernel(const float* __restrict__ k, HT* __restrict__ dst, int Cin, int Cout) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)27 * Cin * Cout) return;
    const int pos = (int)(idx & 63);
    long long r = idx >> 6;
    const int n = (int)(r % Cout); r /= Cout;
    const int chunk = (int)(r % (Cin / 64)); const int tap = (int)(r / (Cin / 64));
    const int ci = chunk * 64 + (((pos >> 3) ^ (n & 7)) << 3) + (pos & 7);
    dst[idx] = HalfOps<HT>::from_float(k[((size_t)tap * Cin + ci) * Cout + n]);
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

template <typename HT, int COUT>
int launch_critic_conv(const void* x, const void* wpack, const float* bias, void* y, const ConvGeom& g, int sm_count, cudaStream_t st) {
    using Cfg = CriticCfg<COUT>;
    EncodeTiledFn enc = encode_fn();
    if (!enc) { rdg_set_error("cuTensorMapEncodeTiled entry point not available"); return RDG_TC_E_DRIVER; }
    if (g.Ci % 64 || g.Co != COUT || g.stride != 2 || g.Wo > 128 || 128 % g.Wo) { rdg_set_error("tc critic conv: unsupported shape"); return RDG_TC_E_SHAPE; }
    CriticConvArgs a{};
    a.B = g.B; a.To = g.To; a.Ho = g.Ho; a.Wo = g.Wo; a.Cin = g.Ci; a.pt = g.pt; a.ph = g.ph; a.pw = g.pw;
    const int plane = g.Ho * g.Wo;
    if (plane >= 128) { a.Hb = 128 / g.Wo; a.Bt = 1; } else { a.Hb = g.Ho; a.Bt = 128 / plane; }
    if (a.Hb < 1 || g.Ho % a.Hb || a.Bt * a.Hb * g.Wo != 128) { rdg_set_error("tc critic conv: output plane does not tile into 128 rows"); return RDG_TC_E_SHAPE; }
    a.n_tiles = ceil_div(g.B, a.Bt) * g.To * (g.Ho / a.Hb);
    a.wpack = wpack; a.bias = bias; a.out = y;
    // input [B,Ti,Hi,Wi,Cin] as (C, W, H, T, B); the box walks W, H (and T) with element stride 2 = the conv stride
    CUtensorMap tmap;
    cuuint64_t gdim[5] = {(cuuint64_t)g.Ci, (cuuint64_t)g.Wi, (cuuint64_t)g.Hi, (cuuint64_t)g.Ti, (cuuint64_t)g.B};
    cuuint64_t gstr[4] = {(cuuint64_t)g.Ci * 2, (cuuint64_t)g.Wi * g.Ci * 2, (cuuint64_t)g.Hi * g.Wi * g.Ci * 2,
                          (cuuint64_t)g.Ti * g.Hi * g.Wi * g.Ci * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)(2 * g.Wo), (cuuint32_t)(2 * a.Hb), 1, (cuuint32_t)a.Bt};
    cuuint32_t estr[5] = {1, 2, 2, 1, 1};
    const CUtensorMapDataType dt = HalfOps<HT>::kFmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = enc(&tmap, dt, 5, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { rdg_set_error("cuTensorMapEncodeTiled (critic) failed: %d", (int)r); return RDG_TC_E_DRIVER; }
    auto kern = tc_critic_conv_kernel<HT, COUT>;
    static bool attr_set = false;
    if (!attr_set) { RDG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem)); attr_set = true; }
    const int grid = a.n_tiles < sm_count ? a.n_tiles : sm_count;
    kern<<<grid, kThreads, Cfg::kSmem, st>>>(tmap, a);
    RDG_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int tc_critic_conv(int half_kind, const void* x, const void* wpack, const float* bias, void* y, const ConvGeom& g, int sm_count, cudaStream_t st) {
    if (g.B <= 0) return 0;
    if (half_kind == RDG_HALF_BF16) {
        if (g.Co == 128) return launch_critic_conv<__nv_bfloat16, 128>(x, wpack, bias, y, g, sm_count, st);
        if (g.Co == 256) return launch_critic_conv<__nv_bfloat16, 256>(x, wpack, bias, y, g, sm_count, st);
    } else {
        if (g.Co == 128) return launch_critic_conv<__half, 128>(x, wpack, bias, y, g, sm_count, st);
        if (g.Co == 256) return launch_critic_conv<__half, 256>(x, wpack, bias, y, g, sm_count, st);
    }
    rdg_set_error("tc critic conv: unsupported Cout %d", g.Co);
    return RDG_TC_E_SHAPE;
}

int pack_critic_weights(int half_kind, const float* k, void* dst, int Cin, int Cout, cudaStream_t st) {
    const long long n = (long long)27 * Cin * Cout;
    if (half_kind == RDG_HALF_BF16) critic_pack_kernel<__nv_bfloat16><<<ceil_div(n, 256), 256, 0, st>>>(k, (__nv_bfloat16*)dst, Cin, Cout);
    else critic_pack_kernel<__half><<<ceil_div(n, 256), 256, 0, st>>>(k, (__half*)dst, Cin, Cout);
    RDG_LAUNCH_CHECK();
    return 0;
}

int critic_first_conv(int half_kind, const float* sample, const float* cond, const float* w, const float* bias, void* out, int B, int nd,
                      int ncond, const ConvGeom& g, cudaStream_t st) {
    if (B <= 0) return 0;
    const long long total = (long long)B * g.To * g.Ho * g.Wo * 8;
    const size_t smem = (size_t)27 * (1 + ncond) * 64 * sizeof(float);
    if (g.Co != 64 || smem > 48 * 1024) { rdg_set_error("critic first conv: unsupported shape"); return RDG_TC_E_SHAPE; }
    if (half_kind == RDG_HALF_BF16)
        critic_first_conv_kernel<__nv_bfloat16><<<ceil_div(total, 256), 256, smem, st>>>(sample, cond, w, bias, (__nv_bfloat16*)out, B, nd, ncond, g.To, g.Ho, g.Wo);
    else
        critic_first_conv_kernel<__half><<<ceil_div(total, 256), 256, smem, st>>>(sample, cond, w, bias, (__half*)out, B, nd, ncond, g.To, g.Ho, g.Wo);
    RDG_LAUNCH_CHECK();
    return 0;
}

int critic_dense_score(int half_kind, const void* h4, const float* w5, const float* b5, float* score, int B, int K, cudaStream_t st) {
    if (B <= 0) return 0;
    if (K & 1) { rdg_set_error("critic dense: odd K"); return RDG_TC_E_SHAPE; }
    if (half_kind == RDG_HALF_BF16) critic_dense_kernel<__nv_bfloat16><<<ceil_div(B, 8), 256, 0, st>>>((const __nv_bfloat16*)h4, w5, b5, score, B, K);
    else critic_dense_kernel<__half><<<ceil_div(B, 8), 256, 0, st>>>((const __half*)h4, w5, b5, score, B, K);
    RDG_LAUNCH_CHECK();
    return 0;
}
