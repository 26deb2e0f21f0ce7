// Probe (test infrastructure): tcgen05.mma (kind::f16, M=128, K=16) issue/execution rate on ONE SM as a function of
// N, operand-view alignment (1024-aligned tile vs row-shifted haloed view with a 9-row group stride) and accumulator
// dependency (same accumulator every MMA vs rotating over several).  Values are irrelevant; only time is measured.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate_probe umma_rate_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tc_mma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF); d |= (uint64_t)(sbo >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
    return d;
}
// mode bits: see host table
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int a_off, int sbo, int ndst, int dst_stride, int reps, int a_alt, long long* cycles, int nissue, int M) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_buf = smem;                // 64 KB of zeros
    uint8_t* b_buf = smem + 64 * 1024;    // 32 KB of zeros
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + 32 * 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int iw = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && iw < nissue) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t bd = make_sdesc(smem_u32(b_buf), 1024);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = tmem_base + (uint32_t)((r % ndst) * dst_stride) + (uint32_t)(iw * 256);
            const uint64_t ad = make_sdesc(smem_u32(a_buf) + a_off + ((a_alt && (r & 1)) ? 23040 : 0) + ((a_alt & 2) ? iw * 23040 : 0), sbo);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_f16(d, ad + 2 * k, bd + 2 * k, idesc, 1u);
        }
        tc_commit(&bars[iw]);
        mbar_wait(&bars[iw], 0);
        long long t1 = clock64();
        cycles[iw] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}
int main() {
    long long* dc; CK(cudaMalloc(&dc, 64));
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    struct Cfg { int N, a_off, sbo, ndst, dst_stride, a_alt; const char* name; int nissue = 1; int M = 128; };
    const int sh = 20 * 128;   // row 20 of a haloed plane
    Cfg cfgs[] = {
        {64, 0, 1024, 1, 0, 0, "N=64  aligned  same D"},
        {64, 0, 1024, 4, 64, 0, "N=64  aligned  4 D rotating"},
        {64, sh, 1152, 1, 0, 0, "N=64  shifted  same D"},
        {64, sh, 1152, 4, 64, 0, "N=64  shifted  4 D rotating"},
        {64, sh, 1152, 2, 128, 1, "N=64  shifted  2 D, A alternating"},
        {128, 0, 1024, 1, 0, 0, "N=128 aligned  same D"},
        {128, 0, 1024, 2, 128, 0, "N=128 aligned  2 D rotating"},
        {128, sh, 1152, 1, 0, 0, "N=128 shifted  same D"},
        {128, sh, 1152, 2, 128, 0, "N=128 shifted  2 D rotating"},
        {256, 0, 1024, 1, 0, 0, "N=256 aligned  same D"},
        {256, sh, 1152, 1, 0, 0, "N=256 shifted  same D"},
        {256, sh, 1152, 2, 256, 0, "N=256 shifted  2 D rotating"},
        {32, 0, 1024, 1, 0, 0, "N=32  aligned  same D"},
        {32, sh, 1152, 4, 32, 0, "N=32  shifted  4 D rotating"},
        {16, 0, 1024, 4, 32, 0, "N=16  aligned  4 D rotating"},
        {64, 128, 1024, 1, 0, 0, "N=64  row+1, sbo 1024 same D"},
        {64, 0, 1152, 1, 0, 0, "N=64  row 0, sbo 1152 same D"},
        {64, sh, 1152, 2, 64, 0, "N=64  shifted 2 issuer warps", 2},
        {128, sh, 1152, 2, 128, 0, "N=128 shifted 2 issuer warps", 2},
        {256, sh, 1152, 1, 0, 0, "N=256 shifted 2 issuer warps", 2},
        {64, sh, 1152, 2, 64, 2, "N=64  2 issuers, different A", 2},
        {128, sh, 1152, 2, 128, 2, "N=128 2 issuers, different A", 2},
        {64, 0, 1024, 2, 64, 2, "N=64  2 issuers, diff A, aligned", 2},
        {64, sh, 1152, 2, 64, 0, "N=64  M=64 shifted", 1, 64},
        {128, sh, 1152, 2, 128, 0, "N=128 M=64 shifted", 1, 64},
        {256, sh, 1152, 1, 0, 0, "N=256 M=64 shifted", 1, 64},
        {192, sh, 1152, 2, 256, 0, "N=192 shifted", 1},
        {160, sh, 1152, 2, 256, 0, "N=160 shifted", 1},
    };
    const int reps = 20000;
    for (auto& c : cfgs) {
        rate_kernel<<<1, 128, 100 * 1024>>>(c.N, c.a_off, c.sbo, c.ndst, c.dst_stride, 200, c.a_alt, dc, c.nissue, c.M);
        rate_kernel<<<1, 128, 100 * 1024>>>(c.N, c.a_off, c.sbo, c.ndst, c.dst_stride, reps, c.a_alt, dc, c.nissue, c.M);
        CK(cudaDeviceSynchronize());
        long long cyc[2]; CK(cudaMemcpy(cyc, dc, 16, cudaMemcpyDeviceToHost));
        printf("%-36s %7.1f clk per MMA per issuer (x%d issuers; math floor %d)\n", c.name, (double)cyc[0] / (reps * 4.0), c.nissue, c.N / 2);
    }
    return 0;
}
