// Probe (test infrastructure, not product): does a tcgen05 K-major SWIZZLE_128B shared-memory descriptor
// accept a start address that is NOT 1024-byte aligned (row offset inside the 8-row swizzle atom) and a
// stride-byte-offset that is not a multiple of 1024, when the tile was written by TMA with SWIZZLE_128B?
// If the XOR is taken from absolute shared-memory address bits on both sides, every (dh, dw) shifted view
// of one haloed activation plane is just a descriptor offset -- the premise of the resident-plane conv kernel.
//
// Layout under test: x[B=2][T][H=8][W=8][C=64] fp16, TMA map with dims (C, W, B, H, T), box (64,10,2,10,1)
// at coordinates (0,-1,0,-1,t): shared image [h' 10][b 2][w' 10][64 ch] with zero halo, 25600 B.
// A view (dh,dw): row r = g*8 + w, g = h*2 + b, at byte ((1+dh)*20 + 1+dw)*128 + g*1280 + w*128.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_shift_probe umma_shift_probe.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t sbo, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

#ifndef WP
#define WP 10
#endif
#ifndef AOFF
#define AOFF 0
#endif
constexpr int kPlane = 10 * 2 * WP * 128;     // box bytes
constexpr int kNW = 128;                      // weight rows (two stacked 64-row tiles -> also tests N=128)

// variant 0: base_offset field = 0; variant 1: base_offset = (start >> 7) & 7
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap, const __half* __restrict__ wsw, float* __restrict__ out, int t,
             int variant, int nmma) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_buf = smem + AOFF;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;   // trailing halo row must be zero
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    uint8_t* b_buf = smem + 48 * 1024;                   // 128 rows x 128 B = 16 KB, pre-swizzled image
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + kNW * 128);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    {
        const uint4* src = reinterpret_cast<const uint4*>(wsw);
        uint4* dst = reinterpret_cast<uint4*>(b_buf);
        for (int i = threadIdx.x; i < kNW * 128 / 16; i += 128) dst[i] = src[i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], kPlane);
        tma_load_5d(a_buf, &tmap, &bars[0], 0, -1, 0, -1, t);
        mbar_wait(&bars[0], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // 9 shifted views, each into its own 64-column accumulator (N=64), then one N=128 MMA for view (0,0) at cols 448? no:
        // columns: view v at v*32?  keep it simple: N=32 per view uses first 32 weight rows -> 9*32 = 288 cols; N=128 test at col 320.
        constexpr uint32_t idesc32 = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t idesc128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int v = 0; v < 9; ++v) {
            const int dh = v / 3 - 1, dw = v % 3 - 1;
            const uint32_t a_addr = smem_u32(a_buf) + ((1 + dh) * 2 * WP + (1 + dw)) * 128;
            const uint32_t bo = variant ? ((a_addr >> 7) & 7) : 0;
            for (int k = 0; k < 4; ++k) {
                const uint64_t ad = make_sdesc(a_addr + k * 32, WP * 128, bo);
                const uint64_t bd = make_sdesc(smem_u32(b_buf) + k * 32, 1024, 0);
                tc_mma_f16(tmem_base + v * 32, ad, bd, idesc32, k > 0);
            }
        }
        {
            const uint32_t a_addr = smem_u32(a_buf) + ((1 + 0) * 2 * WP + (1 + 1)) * 128;      // view (dh=0, dw=+1)
            const uint32_t bo = variant ? ((a_addr >> 7) & 7) : 0;
            for (int rep = 0; rep < nmma; ++rep)
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = make_sdesc(a_addr + k * 32, WP * 128, bo);
                    const uint64_t bd = make_sdesc(smem_u32(b_buf) + k * 32, 1024, 0);
                    tc_mma_f16(tmem_base + 320, ad, bd, idesc128, (rep | k) > 0);
                }
        }
        tc_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int r = warp * 32 + lane;
    for (int c0 = 0; c0 < 448; c0 += 32) {
        uint32_t v[32];
        tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int j = 0; j < 32; ++j) out[(size_t)r * 448 + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int B = 2, T = 3, H = 8, W = 8, C = 64;
    std::vector<__half> x((size_t)B * T * H * W * C);
    std::vector<float> xf(x.size());
    srand(1);
    for (size_t i = 0; i < x.size(); ++i) { int v = rand() % 9 - 4; x[i] = __float2half((float)v); xf[i] = (float)v; }
    std::vector<float> wf((size_t)kNW * 64);
    std::vector<__half> wsw((size_t)kNW * 64);
    for (int n = 0; n < kNW; ++n)
        for (int k = 0; k < 64; ++k) {
            int v = rand() % 7 - 3;
            wf[n * 64 + k] = (float)v;
            const int j = (k >> 3) ^ (n & 7), e = k & 7;      // chunk k>>3 of row n lives at chunk (k>>3)^(n&7)
            wsw[n * 64 + j * 8 + e] = __float2half((float)v);
        }
    __half *dx, *dw; float* dout;
    CK(cudaMalloc(&dx, x.size() * 2)); CK(cudaMalloc(&dw, wsw.size() * 2)); CK(cudaMalloc(&dout, 128 * 448 * 4));
    CK(cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, wsw.data(), wsw.size() * 2, cudaMemcpyHostToDevice));

    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    CUtensorMap tmap;
    // dims (C, W, B, H, T): strides of dims 1..4 in bytes
    cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)B, (cuuint64_t)H, (cuuint64_t)T};
    cuuint64_t gstr[4] = {(cuuint64_t)C * 2, (cuuint64_t)T * H * W * C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[5] = {64, WP, 2, 10, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, dx, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d (non-monotonic strides rejected?)\n", (int)r); return 1; }
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));

    int ok_any = 0;
    for (int variant = 0; variant < 2; ++variant) {
        for (int t = 0; t < T; t += 2) {
            CK(cudaMemset(dout, 0xff, 128 * 448 * 4));
            probe_kernel<<<1, 128, 80 * 1024>>>(tmap, dw, dout, t, variant, 1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("variant %d: kernel failed: %s\n", variant, cudaGetErrorString(e)); return 3; }
            std::vector<float> out(128 * 448);
            CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
            int bad_views = 0;
            for (int v = 0; v < 10; ++v) {
                const int dh = v < 9 ? v / 3 - 1 : 0, dwv = v < 9 ? v % 3 - 1 : 1;
                const int N = v < 9 ? 32 : 128, col0 = v < 9 ? v * 32 : 320;
                int bad = 0;
                for (int row = 0; row < 128; ++row) {
                    const int g = row / 8, w = row % 8, h = g / 2, b = g % 2;
                    const int hh = h + dh, ww = w + dwv;
                    for (int n = 0; n < N; ++n) {
                        float ref = 0.f;
                        if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                            for (int k = 0; k < 64; ++k) ref += xf[((((size_t)b * T + t) * H + hh) * W + ww) * C + k] * wf[n * 64 + k];
                        if (out[(size_t)row * 448 + col0 + n] != ref) ++bad;
                    }
                }
                printf("variant %d t %d view %d (dh %d dw %d N %d): %s (%d mismatches)\n", variant, t, v, dh, dwv, N, bad ? "MISMATCH" : "ok", bad);
                bad_views += bad != 0;
            }
            if (!bad_views) { ok_any |= 1 << variant; }
        }
    }
    printf("RESULT: base_offset=0 %s ; base_offset=(addr>>7)&7 %s\n", (ok_any & 1) ? "WORKS" : "fails", (ok_any & 2) ? "WORKS" : "fails");

    // timing: N=128 MMA on a shifted view, many repetitions, to see whether unaligned views cost extra cycles
    for (int variant = 0; variant < 2; ++variant) {
        if (!(ok_any & (1 << variant))) continue;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        const int nm = 20000;
        probe_kernel<<<1, 128, 80 * 1024>>>(tmap, dw, dout, 0, variant, 100);
        CK(cudaEventRecord(e0));
        probe_kernel<<<1, 128, 80 * 1024>>>(tmap, dw, dout, 0, variant, nm);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("variant %d: %d x4 N=128 K=16 MMAs on the shifted view: %.3f ms -> %.1f ns per MMA\n", variant, nm, ms, ms * 1e6 / (nm * 4.0));
    }
    return ok_any ? 0 : 4;
}
