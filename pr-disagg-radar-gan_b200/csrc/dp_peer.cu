// Data-parallel gradient exchange over NVLink peer memory (one node, one process per GPU) -- SURVEY 8e, row e2.
//
// The reference trains on one GPU; the data-parallel form averages the two flat FP32 gradient buffers over the ranks before the
// shared Adam update (gan_train_cwgangp_pixelnorm.py:385, :412; batch-mean losses :215-216, so N ranks x batch B == one rank x
// batch N*B).  NCCL does that with one all-reduce per optimizer step, but an NCCL call cannot sit inside the captured iteration
// (capturing it hung), which forces one graph per step phase with eager launches in between.  This file is the exchange as the
// library's own kernels, so that a whole data-parallel iteration is ONE CUDA graph:
//
//   * every rank exports its two gradient buffers and a small flag array as CUDA IPC handles; rdg_peer_connect maps the peers'
//     (NVSwitch: every GPU reaches every peer at full rate, no ring order to respect);
//   * rdg_peer_allreduce = barrier, one kernel, barrier.  The kernel is a two-shot all-reduce in place: rank r owns the r-th
//     slice of the buffer, loads that slice from every rank (its own included) in rank order 0..N-1 -- so all ranks would form
//     bit-identical sums, and each element is summed exactly once, by its owner -- and stores the sum into the same slice of every
//     rank's buffer.  Per rank that moves (N-1)/N of the buffer in and out over NVLink instead of N-1 whole buffers;
//   * the barriers are flag words in peer-visible memory: rank r stores the (monotonic, device-resident) epoch into slot r of every
//     rank's flag array with a system-scope release and spins on its own array with system-scope acquires.  The first barrier
//     publishes the gradients (they were written by earlier kernels of the same stream), the second keeps a rank from zeroing
//     its buffer for the next step while a peer still reads it.  A spin that lasts 60 s sets an error word and gives up
//     (a rank that died must not hang the others' GPUs); rdg_peer_status reports it.
// The 1/N scale stays fused in the Adam kernel (grad_scale), exactly as with NCCL.
#include <cstdint>
#include <cstring>

#include "../../include/rdg_b200.h"
#include "ctx.h"
#include "rdg_common.cuh"

namespace {

constexpr int kMaxRanks = 8;
constexpr int kErrSlot = 16;                 // flags[16]: set when a barrier timed out
constexpr int kFlagWords = 32;
constexpr unsigned long long kBarrierTimeoutNs = 60ull * 1000000000ull;   // ranks may be seconds apart in their first iteration (allocations)
constexpr int kHandleBytes = 3 * 64;         // generator gradients, critic gradients, flags

#define TRY(x) do { int r_ = (x); if (r_) return r_; } while (0)

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// one warp; lane r talks to rank r
__global__ void peer_barrier_kernel(unsigned* const* __restrict__ peer_flags, unsigned* __restrict__ my_flags, unsigned* __restrict__ epoch,
                                    int world, int rank) {
    const unsigned k = *epoch + 1;
    __syncwarp();
    const int r = threadIdx.x;
    if (r < world) {
        __threadfence_system();
        st_release_sys(peer_flags[r] + rank, k);
        const unsigned long long t0 = globaltimer_ns();
        while ((int)(ld_acquire_sys(my_flags + r) - k) < 0) {
            if (globaltimer_ns() - t0 > kBarrierTimeoutNs) { my_flags[kErrSlot] = 1u; break; }
            __nanosleep(64);
        }
    }
    __syncwarp();
    if (threadIdx.x == 0) *epoch = k;
}

// in-place two-shot all-reduce (sum) of n4 float4s; this rank owns [lo4, hi4)
__global__ void __launch_bounds__(256) peer_reduce_kernel(float* const* __restrict__ peer_g, long long lo4, long long hi4, int world) {
    const long long i = lo4 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi4) return;
    float4 s = __ldcg(reinterpret_cast<const float4*>(peer_g[0]) + i);
    for (int r = 1; r < world; ++r) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(peer_g[r]) + i);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    for (int r = 0; r < world; ++r) __stcg(reinterpret_cast<float4*>(peer_g[r]) + i, s);
}

int peer_free(rdg_ctx* c) {
    RdgPeer& p = c->peer;
    for (int r = 0; r < kMaxRanks; ++r)
        for (int j = 0; j < 3; ++j)
            if (p.mapped[r][j]) { cudaIpcCloseMemHandle(p.mapped[r][j]); p.mapped[r][j] = nullptr; }
    cudaFree(p.d_grads[0]); cudaFree(p.d_grads[1]); cudaFree(p.d_flags); cudaFree(p.epoch);
    p.d_grads[0] = p.d_grads[1] = nullptr; p.d_flags = nullptr; p.epoch = nullptr;
    p.world = 0;
    return 0;
}

}  // namespace

void rdg_peer_destroy(rdg_ctx* c) {
    peer_free(c);
    cudaFree(c->peer.flags);
    c->peer.flags = nullptr;
}

extern "C" int rdg_peer_handle_bytes(void) { return kHandleBytes; }

extern "C" int rdg_peer_export(rdg_ctx* c, unsigned char* handles) {
    if (!c || !handles) { rdg_set_error("rdg_peer_export: bad arguments"); return RDG_E_BADARG; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    RDG_CUDA(cudaSetDevice(c->device));
    RdgPeer& p = c->peer;
    if (!p.flags) {
        RDG_CUDA(cudaMalloc(&p.flags, kFlagWords * sizeof(unsigned)));
        RDG_CUDA(cudaMemset(p.flags, 0, kFlagWords * sizeof(unsigned)));
    }
    for (int which = 0; which < 2; ++which) {
        float* g = nullptr; size_t n = 0;
        TRY(rdg_grad_buffer(c, which, &g, &n));
        cudaIpcMemHandle_t h;
        RDG_CUDA(cudaIpcGetMemHandle(&h, g));
        memcpy(handles + 64 * which, &h, 64);
    }
    cudaIpcMemHandle_t h;
    RDG_CUDA(cudaIpcGetMemHandle(&h, p.flags));
    memcpy(handles + 128, &h, 64);
    return 0;
}

extern "C" int rdg_peer_connect(rdg_ctx* c, int world, int rank, const unsigned char* all_handles) {
    if (!c || !all_handles || world < 2 || world > kMaxRanks || rank < 0 || rank >= world) {
        rdg_set_error("rdg_peer_connect: need 2 <= world <= %d, 0 <= rank < world and the gathered handles", kMaxRanks);
        return RDG_E_BADARG;
    }
    RDG_CUDA(cudaSetDevice(c->device));
    RdgPeer& p = c->peer;
    if (!p.flags) { rdg_set_error("rdg_peer_connect: call rdg_peer_export first"); return RDG_E_BADARG; }
    RDG_CUDA(cudaDeviceSynchronize());
    peer_free(c);
    float* g_of[2][kMaxRanks]; unsigned* f_of[kMaxRanks];
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            for (int which = 0; which < 2; ++which) { size_t n; TRY(rdg_grad_buffer(c, which, &g_of[which][r], &n)); }
            f_of[r] = p.flags;
            continue;
        }
        for (int j = 0; j < 3; ++j) {
            cudaIpcMemHandle_t h;
            memcpy(&h, all_handles + (size_t)r * kHandleBytes + 64 * j, 64);
            void* ptr = nullptr;
            RDG_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            p.mapped[r][j] = ptr;
            if (j < 2) g_of[j][r] = static_cast<float*>(ptr); else f_of[r] = static_cast<unsigned*>(ptr);
        }
    }
    for (int which = 0; which < 2; ++which) {
        RDG_CUDA(cudaMalloc(&p.d_grads[which], kMaxRanks * sizeof(float*)));
        RDG_CUDA(cudaMemcpy(p.d_grads[which], g_of[which], world * sizeof(float*), cudaMemcpyHostToDevice));
    }
    RDG_CUDA(cudaMalloc(&p.d_flags, kMaxRanks * sizeof(unsigned*)));
    RDG_CUDA(cudaMemcpy(p.d_flags, f_of, world * sizeof(unsigned*), cudaMemcpyHostToDevice));
    RDG_CUDA(cudaMalloc(&p.epoch, sizeof(unsigned)));
    // the flag words keep counting across reconnects: start this rank's epoch where its own slot stands (all ranks make the same
    // number of barrier calls, so the slots of a rank's array agree whenever no exchange is in flight)
    RDG_CUDA(cudaMemcpy(p.epoch, p.flags + rank, sizeof(unsigned), cudaMemcpyDeviceToDevice));
    p.world = world; p.rank = rank;
    return 0;
}

extern "C" int rdg_peer_disconnect(rdg_ctx* c) {
    if (!c) return RDG_E_BADARG;
    RDG_CUDA(cudaSetDevice(c->device));
    RDG_CUDA(cudaDeviceSynchronize());
    return peer_free(c);
}

extern "C" int rdg_peer_allreduce(rdg_ctx* c, int which, void* stream) {
    if (!c || which < 0 || which > 1) { rdg_set_error("rdg_peer_allreduce: bad arguments"); return RDG_E_BADARG; }
    RdgPeer& p = c->peer;
    if (p.world < 2) { rdg_set_error("rdg_peer_allreduce: not connected (rdg_peer_connect)"); return RDG_E_BADARG; }
    RDG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long long n4 = (long long)((which == 0 ? c->g_total : c->c_total) / 4);     // flat buffers are padded to multiples of 4
    const long long per = (n4 + p.world - 1) / p.world;
    const long long lo = per * p.rank < n4 ? per * p.rank : n4, hi = lo + per < n4 ? lo + per : n4;
    peer_barrier_kernel<<<1, 32, 0, st>>>(p.d_flags, p.flags, p.epoch, p.world, p.rank);
    RDG_LAUNCH_CHECK();
    if (hi > lo) {
        peer_reduce_kernel<<<(unsigned)((hi - lo + 255) / 256), 256, 0, st>>>(p.d_grads[which], lo, hi, p.world);
        RDG_LAUNCH_CHECK();
    }
    peer_barrier_kernel<<<1, 32, 0, st>>>(p.d_flags, p.flags, p.epoch, p.world, p.rank);
    RDG_LAUNCH_CHECK();
    return 0;
}

// timeouts: 1 if a barrier of this rank ever gave up waiting for a peer (the sums of that exchange are then not the global sums)
extern "C" int rdg_peer_status(rdg_ctx* c, int* world, int* timeouts) {
    if (!c) return RDG_E_BADARG;
    RDG_CUDA(cudaSetDevice(c->device));
    if (world) *world = c->peer.world;
    if (timeouts) {
        unsigned e = 0;
        if (c->peer.flags) RDG_CUDA(cudaMemcpy(&e, c->peer.flags + kErrSlot, sizeof(e), cudaMemcpyDeviceToHost));
        *timeouts = (int)e;
    }
    return 0;
}
