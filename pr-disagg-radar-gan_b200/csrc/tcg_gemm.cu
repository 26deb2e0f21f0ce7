// Training-path implicit GEMMs on tcgen05 (kind::tf32, FP32 accumulate in TMEM) -- see tcg.h.
//
// Two kernels cover every contraction of the training step:
//
//  tcg_rowgemm_kernel   D[row, n] = sum_{tap, c} SRC[pos(row) (+) tap][c] * W[tap][n][c]
//      rows = output positions (forward convs: critic stride-2 'same'/'valid', Dense, the 8 phase convs of an upsample-folded
//      block) or input positions (backward-data: stride-2 transposed convs by parity class, the folded block's backward straight
//      to the low-res grid).  All of them are the same thing after the host has split the row set into CLASSES (parity classes /
//      output phases): inside a class the source coordinate is an affine map a*t + c_tap of the class-local coordinate, the set
//      of contributing taps is the same for every row, and the output coordinate is o + os*t.  128 rows per CTA (UMMA M = 128),
//      N <= 256 accumulator columns, K streamed in 32-element (128-byte) blocks through a shared-memory ring.
//
//  tcg_filtergrad_kernel   dW[tap][ci][co] += sum_pos X[pos (+) tap][ci] * DY[pos][co]
//      the contraction runs over positions, which are NOT contiguous in the channels-last tensors: the producer warps read 16-byte
//      channel vectors per position and scatter them transposed into the K-major operand tiles (lane = position: conflict-free).
//
// Both: 8 producer warps gather operands global -> registers (cvt.rna.tf32) -> 128-byte-swizzled K-major shared memory
// (fence.proxy.async, mbarrier full/empty ring), one elected thread of warp 8 issues tcgen05.mma, the producer warps turn into the
// epilogue (tcgen05.ld 32x32b: thread = accumulator row).  Split-K: deterministic partial slices + splitk reduction (row GEMM),
// vector red.global.add (filter gradients, which accumulate into the shared gradient buffer anyway).
#include "tc_ptx.cuh"
#include "tcg.h"
#include "gen_tc.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace {
using namespace rdg_tc;

constexpr int TCG_THREADS = 288;
constexpr int MAX_STAGES = 8;

struct TcgTap { int8_t ct, ch, cw, pad_; int32_t widx; };
struct TcgClass { int Tc, Hc, Wc, rows, tile_begin, ot, oh, ow, tap_begin, tap_count; };

struct TcgRowArgs {
    const float* src; const float* w; float* out; float* part;
    const float* bias; float* pre; const float* mask;
    float mask_scale; int act;
    int a, os;                // source coordinate scale, output coordinate step
    int Ts, Hs, Ws, Cs;       // source tensor (channels-last)
    int To, Ho, Wo, Nt;       // output tensor
    int N;                    // accumulator columns of one CTA (divides Nt)
    int Kc, kchunks;          // contraction channels per tap, 32-element blocks per tap
    int wrow;                 // weight elements between consecutive n
    long long wtap;           // weight elements per tap block
    long long out_elems;      // elements of the output tensor (stride between split-K slices)
    int nclass, nslice, nstages;
    int ksmall;               // > 0: few-channel input, k = tap * Cs + channel runs over ksmall = taps * Cs valid elements (Kc = padded)
    const float* src2;        // ksmall only, != null: channel 0 comes from `src` [B,Ts,Hs,Ws] (1 channel) and channels 1.. from src2
                              // [B,Hs,Ws,Cs-1] broadcast over t: the critic's sample / condition concat (:275-282) fused into the gather
    void* out16; int half_kind;   // != null: the result is stored as 16-bit (RDG_HALF_BF16 / RDG_HALF_FP16) instead of FP32 `out`
    TcgClass cls[8];
    TcgTap taps[64];
};

struct TcgFTap { int8_t xt, xh, xw, yt, yh, yw; int16_t widx; };
struct TcgFilterArgs {
    const float* x; const float* dy; float* dw;
    int Tc, Hc, Wc, rows;     // iteration grid per sample; rows = B*Tc*Hc*Wc positions
    int ax, Tx, Hx, Wx, Ci;
    int ay, Ty, Hy, Wy, Co;
    int Mt, N, ksplit, nstages;
    int ksmall;               // > 0: few-channel input: A rows = (tap, channel), ksmall = taps * Ci of them, one launch for all taps
    unsigned long long mW, mH, mT;   // fast_div multipliers of Wc, Hc, Tc
    int cpa;                  // 16-byte channel chunks per position of this launch's A tile: min(Ci, 128) / 4, a power of two
    long long wblk;           // dw elements per tap block (Ci*Co)
    TcgFTap taps[64];
};

// x / d for 0 <= x < 2^24, 1 <= d < 256 with the host-computed multiplier ceil(2^40 / d): exact (x d^2 < 2^40), 4 instructions
// instead of the ~35 of a runtime integer division (the producers decompose two positions per k-block)
__device__ __forceinline__ int fast_div(int x, unsigned long long m) { return (int)(((unsigned long long)(unsigned)x * m) >> 40); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void sts4_tf32(uint8_t* dst, float4 v) {
    float4 r = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    *reinterpret_cast<float4*>(dst) = r;
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ uint32_t tmem_cols_for(int N) { return N <= 32 ? 32u : N <= 64 ? 64u : N <= 128 ? 128u : 256u; }
__device__ __forceinline__ uint32_t idesc_tf32(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);   // F32 accum, tf32 x tf32, K-major, M = 128
}

// MMA issuer loop shared by both kernels.  Stage layout: A (128 rows) at +0, B (32 NB rows) at +16384; SPLIT (3xTF32): the low parts
// A_lo / B_lo follow in the second half of the stage and every k step issues hi*hi + lo*hi + hi*lo.
template <bool SPLIT>
__device__ __forceinline__ void mma_issue_loop(uint8_t* smem, uint32_t stage_bytes, int nst, int nkb, int N, uint32_t tmem,
                                               uint64_t* full_bar, uint64_t* empty_bar, uint64_t* acc_bar) {
    const uint32_t idesc = idesc_tf32(N);
    const uint32_t half = SPLIT ? stage_bytes / 2 : 0;
    for (int i = 0; i < nkb; ++i) {
        const int s = i % nst;
        mbar_wait(&full_bar[s], (uint32_t)(i / nst) & 1u);
        if (!SPLIT) fence_proxy_async_smem();      // cp.async producers: the landed (generic-proxy) writes -> async proxy
        tc_fence_after();
        if (elect_one()) {
            const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
            const uint64_t ad = make_sdesc(base), bd = make_sdesc(base + 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                tc_mma_tf32(tmem, ad + 2 * k, bd + 2 * k, idesc, (i | k) ? 1u : 0u);
                if (SPLIT) {
                    const uint64_t al = make_sdesc(base + half), bl = make_sdesc(base + half + 16384);
                    tc_mma_tf32(tmem, al + 2 * k, bd + 2 * k, idesc, 1u);
                    tc_mma_tf32(tmem, ad + 2 * k, bl + 2 * k, idesc, 1u);
                }
            }
            tc_commit(&empty_bar[s]);
            if (i == nkb - 1) tc_commit(acc_bar);
        }
        __syncwarp();
    }
}

template <bool SPLIT>
__device__ __forceinline__ void sts4_op(uint8_t* dst, uint32_t lo_off, float4 v) {
    const float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    *reinterpret_cast<float4*>(dst) = h;
    if (SPLIT) *reinterpret_cast<float4*>(dst + lo_off) = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
}

// ------------------------------------------------------------------------------------------------ row GEMM
// NB = N / 32 (32-row slabs of the B tile); SPLIT = 3xTF32 (forward passes: FP32-grade pre-activations keep the LeakyReLU slope
// pattern of the FP64 oracle); two CTAs per SM where the register / shared-memory budget allows.
template <int NB, bool SPLIT>
__global__ void __launch_bounds__(TCG_THREADS, (NB * (SPLIT ? 2 : 1) <= 4) ? 2 : 1) tcg_rowgemm_kernel(const __grid_constant__ TcgRowArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], acc_bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int N = NB * 32;
    const int nst = p.nstages;
    constexpr uint32_t stage_bytes = (16384u + (uint32_t)N * 128u) * (SPLIT ? 2u : 1u);

    int ci = 0;
    for (int i = 1; i < p.nclass; ++i)
        if ((int)blockIdx.x >= p.cls[i].tile_begin) ci = i;
    const TcgClass cl = p.cls[ci];
    const int row0 = ((int)blockIdx.x - cl.tile_begin) * 128;
    const int n0 = blockIdx.y * N;
    const int nkb_all = p.ksmall ? p.kchunks : cl.tap_count * p.kchunks;
    const int kb_lo = (int)((long long)blockIdx.z * nkb_all / p.nslice);
    const int nkb = (int)((long long)(blockIdx.z + 1) * nkb_all / p.nslice) - kb_lo;

    if (tid == 0) {
        for (int s = 0; s < nst; ++s) { mbar_init(&full_bar[s], SPLIT ? 8 : 256); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc(&tmem_slot, tmem_cols_for(N));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 8) {
        mma_issue_loop<SPLIT>(smem, stage_bytes, nst, nkb, N, tmem, full_bar, empty_bar, &acc_bar);
    } else {
        // ---- producers: thread = (row sub-index rsub, 16-byte chunk) of every 32-row slab of the A and B tiles
        const int chunk = tid & 7, rsub = tid >> 3;
        int rbase[4], rcoord[4], rb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = row0 + rsub + 32 * j;
            rbase[j] = -1; rcoord[j] = 0; rb[j] = 0;
            if (r < cl.rows) {
                int q = r;
                const int w2 = q % cl.Wc; q /= cl.Wc;
                const int h2 = q % cl.Hc; q /= cl.Hc;
                const int t2 = q % cl.Tc; const int b = q / cl.Tc;
                const int ts = p.a * t2, hs = p.a * h2, ws = p.a * w2;
                rbase[j] = (((b * p.Ts + ts) * p.Hs + hs) * p.Ws + ws) * p.Cs;
                rcoord[j] = ts | (hs << 10) | (ws << 20);
                rb[j] = b;
            }
        }
        const uint32_t sw_off = (uint32_t)(rsub >> 3) * 1024u + (uint32_t)(rsub & 7) * 128u + (uint32_t)((chunk ^ (rsub & 7)) << 4);
        auto load = [&](int i, float4 (&va)[4], float4 (&vb)[NB]) {
            const int kb = kb_lo + i;
            int te = 0, kc = kb;
            if (!p.ksmall) { te = kb / p.kchunks; kc = kb - te * p.kchunks; }
            const TcgTap tp = p.taps[cl.tap_begin + te];
            const int k0 = kc * 32 + chunk * 4;
            const bool kval = k0 < p.Kc;
            const float* wb = p.w + (long long)(p.ksmall ? 0 : tp.widx) * p.wtap + (long long)(n0 + rsub) * p.wrow + k0;
#pragma unroll
            for (int j = 0; j < NB; ++j) vb[j] = kval ? ldg4(wb + (long long)(32 * j) * p.wrow) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (!p.ksmall) {
                const int doff = ((tp.ct * p.Hs + tp.ch) * p.Ws + tp.cw) * p.Cs + k0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ts = (rcoord[j] & 1023) + tp.ct, hs = ((rcoord[j] >> 10) & 1023) + tp.ch, ws = (rcoord[j] >> 20) + tp.cw;
                    const bool ok = kval && rbase[j] >= 0 && (unsigned)ts < (unsigned)p.Ts && (unsigned)hs < (unsigned)p.Hs &&
                                    (unsigned)ws < (unsigned)p.Ws;
                    va[j] = ok ? ldg4(p.src + rbase[j] + doff) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                // few input channels (the critic's first conv, Cs = 1 + ncond): k = tap * Cs + channel, element-wise gather
                int eo[4], ect[4], ech[4], ecw[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = k0 + u;
                    eo[u] = -1; ect[u] = ech[u] = ecw[u] = 0;
                    if (k < p.ksmall) {
                        const int tap = k / p.Cs, c = k - tap * p.Cs;
                        const TcgTap tq = p.taps[cl.tap_begin + tap];
                        ect[u] = tq.ct; ech[u] = tq.ch; ecw[u] = tq.cw;
                        eo[u] = ((tq.ct * p.Hs + tq.ch) * p.Ws + tq.cw) * p.Cs + c;
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float e[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int ts = (rcoord[j] & 1023) + ect[u], hs = ((rcoord[j] >> 10) & 1023) + ech[u], ws = (rcoord[j] >> 20) + ecw[u];
                        const bool ok = eo[u] != -1 && rbase[j] >= 0 && (unsigned)ts < (unsigned)p.Ts && (unsigned)hs < (unsigned)p.Hs &&
                                        (unsigned)ws < (unsigned)p.Ws;
                        e[u] = ok ? __ldg(p.src + rbase[j] + eo[u]) : 0.f;
                    }
                    va[j] = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
        };
        auto store = [&](int i, const float4 (&va)[4], const float4 (&vb)[NB]) {
            const int s = i % nst;
            if (i >= nst) mbar_wait(&empty_bar[s], ((uint32_t)(i / nst) & 1u) ^ 1u);
            uint8_t* sa = smem + (size_t)s * stage_bytes + sw_off;
#pragma unroll
            for (int j = 0; j < 4; ++j) sts4_op<SPLIT>(sa + j * 4096, stage_bytes / 2, va[j]);
#pragma unroll
            for (int j = 0; j < NB; ++j) sts4_op<SPLIT>(sa + 16384 + j * 4096, stage_bytes / 2, vb[j]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[s]);
        };
        if (SPLIT) {
            // register path (operands are split into tf32 high + low parts on the way): software pipeline, the loads of k-block
            // i+1 are in flight while k-block i is rounded and stored
            float4 a0[4], b0[NB], a1[4], b1[NB];
            if (nkb > 0) load(0, a0, b0);
            for (int i = 0; i < nkb; i += 2) {
                if (i + 1 < nkb) load(i + 1, a1, b1);
                store(i, a0, b0);
                if (i + 1 < nkb) {
                    if (i + 2 < nkb) load(i + 2, a0, b0);
                    store(i + 1, a1, b1);
                }
            }
        } else {
            // asynchronous path: cp.async straight into the swizzled tiles (the tensor core truncates FP32 bit patterns to tf32),
            // completion arrives on the stage's full barrier; the producers only wait for free stages, so `nstages` k-blocks
            // are in flight per CTA
            for (int i = 0; i < nkb; ++i) {
                const int kb = kb_lo + i;
                int te = 0, kc = kb;
                if (!p.ksmall) { te = kb / p.kchunks; kc = kb - te * p.kchunks; }
                const TcgTap tp = p.taps[cl.tap_begin + te];
                const int k0 = kc * 32 + chunk * 4;
                const bool kval = k0 < p.Kc;
                const int s = i % nst;
                if (i >= nst) mbar_wait(&empty_bar[s], ((uint32_t)(i / nst) & 1u) ^ 1u);
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes + sw_off);
                const float* wb = p.w + (long long)(p.ksmall ? 0 : tp.widx) * p.wtap + (long long)(n0 + rsub) * p.wrow + (kval ? k0 : 0);
#pragma unroll
                for (int j = 0; j < NB; ++j) cp_async16(sa + 16384 + j * 4096, wb + (long long)(32 * j) * p.wrow, kval ? 16u : 0u);
                if (!p.ksmall) {
                    const int doff = ((tp.ct * p.Hs + tp.ch) * p.Ws + tp.cw) * p.Cs + k0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int ts = (rcoord[j] & 1023) + tp.ct, hs = ((rcoord[j] >> 10) & 1023) + tp.ch, ws = (rcoord[j] >> 20) + tp.cw;
                        const bool ok = kval && rbase[j] >= 0 && (unsigned)ts < (unsigned)p.Ts && (unsigned)hs < (unsigned)p.Hs &&
                                        (unsigned)ws < (unsigned)p.Ws;
                        cp_async16(sa + j * 4096, ok ? p.src + rbase[j] + doff : p.src, ok ? 16u : 0u);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = k0 + u;
                        int eo = -1, ect = 0, ech = 0, ecw = 0, ec = 0;
                        if (k < p.ksmall) {
                            const int tap = k / p.Cs, c = k - tap * p.Cs;
                            const TcgTap tq = p.taps[cl.tap_begin + tap];
                            ect = tq.ct; ech = tq.ch; ecw = tq.cw; ec = c;
                            eo = ((tq.ct * p.Hs + tq.ch) * p.Ws + tq.cw) * p.Cs + c;
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int ts = (rcoord[j] & 1023) + ect, hs = ((rcoord[j] >> 10) & 1023) + ech, ws = (rcoord[j] >> 20) + ecw;
                            const bool ok = eo != -1 && rbase[j] >= 0 && (unsigned)ts < (unsigned)p.Ts && (unsigned)hs < (unsigned)p.Hs &&
                                            (unsigned)ws < (unsigned)p.Ws;
                            const float* g = p.src;
                            if (ok) {
                                if (!p.src2) g = p.src + rbase[j] + eo;
                                else if (ec == 0) g = p.src + ((rb[j] * p.Ts + ts) * p.Hs + hs) * p.Ws + ws;
                                else g = p.src2 + ((rb[j] * p.Hs + hs) * p.Ws + ws) * (p.Cs - 1) + (ec - 1);
                            }
                            cp_async4(sa + j * 4096 + 4 * u, g, ok ? 4u : 0u);
                        }
                    }
                }
                cp_async_mbar_arrive_noinc(&full_bar[s]);
            }
        }
        // ---- epilogue: warp = (lane quarter q, column half)
        if (nkb > 0) {        // an empty K slice (class with fewer taps than slices) contributes zeros and never touches TMEM
            mbar_wait(&acc_bar, 0);
            tc_fence_after();
        }
        const int q = warp & 3, half = warp >> 2;
        constexpr int cph = N >= 64 ? N / 2 : 32;
        const int c_lo = half * cph, c_hi = min(N, c_lo + cph);
        const int r = row0 + q * 32 + lane;
        long long off = -1;
        if (r < cl.rows) {
            int qq = r;
            const int w2 = qq % cl.Wc; qq /= cl.Wc;
            const int h2 = qq % cl.Hc; qq /= cl.Hc;
            const int t2 = qq % cl.Tc; const int b = qq / cl.Tc;
            off = ((((long long)b * p.To + cl.ot + p.os * t2) * p.Ho + cl.oh + p.os * h2) * p.Wo + cl.ow + p.os * w2) * p.Nt + n0;
        }
        for (int c = c_lo; c < c_hi; c += 32) {
            uint32_t v[32];
            if (nkb > 0) tc_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            if (off < 0) continue;
            if (p.part) {
                float* dst = p.part + (long long)blockIdx.z * p.out_elems + off + c;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                      __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float x[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        x[e] = __uint_as_float(v[j + e]);
                        if (p.bias) x[e] += __ldg(p.bias + n0 + c + j + e);
                    }
                    if (p.act == ACT_LRELU_BWD) {          // fused LeakyReLU backward: the cotangent of the pre-activation
                        const float4 a = ldg4(p.pre + off + c + j);
                        x[0] *= a.x > 0.f ? 1.f : 0.2f; x[1] *= a.y > 0.f ? 1.f : 0.2f;
                        x[2] *= a.z > 0.f ? 1.f : 0.2f; x[3] *= a.w > 0.f ? 1.f : 0.2f;
                    } else if (p.pre) {
                        *reinterpret_cast<float4*>(p.pre + off + c + j) = make_float4(x[0], x[1], x[2], x[3]);
                    }
                    if (p.act == ACT_LRELU) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) x[e] = x[e] > 0.f ? x[e] : 0.2f * x[e];
                    }
                    if (p.mask) {
                        const float4 m = ldg4(p.mask + off + c + j);
                        x[0] *= m.x * p.mask_scale; x[1] *= m.y * p.mask_scale; x[2] *= m.z * p.mask_scale; x[3] *= m.w * p.mask_scale;
                    }
                    if (p.out16) {
                        uint2 h;
                        if (p.half_kind == RDG_HALF_BF16) { h.x = HalfOps<__nv_bfloat16>::pack(x[0], x[1]); h.y = HalfOps<__nv_bfloat16>::pack(x[2], x[3]); }
                        else { h.x = HalfOps<__half>::pack(x[0], x[1]); h.y = HalfOps<__half>::pack(x[2], x[3]); }
                        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out16) + off + c + j) = h;
                    } else {
                        *reinterpret_cast<float4*>(p.out + off + c + j) = make_float4(x[0], x[1], x[2], x[3]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem, tmem_cols_for(N)); }
}

// out = act(sum of slices + bias) [* mask]; same contract as the SIMT split-K epilogue.  G thread groups of a block share the
// slices of 256 / G outputs (small outputs with many slices: the deep layers at batch 32 have 16-96 blocks' worth of outputs and
// 27-54 slices, one thread walking them all is a 10-15 us latency chain); the sum order is fixed, so results are reproducible.
template <int G>
__global__ void __launch_bounds__(256) tcg_splitk_epilogue_kernel(const float* __restrict__ part, int nslice, long long MN4, int N,
                                                                  const float* __restrict__ bias, float* __restrict__ y, int act,
                                                                  const float* __restrict__ mask, float mask_scale, float* __restrict__ pre) {
    constexpr int W = 256 / G;
    const int tx = threadIdx.x % W, ty = threadIdx.x / W;
    const long long i = (long long)blockIdx.x * W + tx;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < MN4) {
#pragma unroll 4
        for (int z = ty; z < nslice; z += G) {
            const float4 t = reinterpret_cast<const float4*>(part)[(long long)z * MN4 + i];
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
    }
    if (G > 1) {
        __shared__ float4 red[256];
        red[threadIdx.x] = s;
        __syncthreads();
        if (ty) return;
#pragma unroll
        for (int g = 1; g < G; ++g) {
            const float4 t = red[g * W + tx];
            s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
    }
    if (i >= MN4) return;
    if (bias) {
        const float4 b = ldg4(bias + (int)((i * 4) % N));
        s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
    }
    if (act == ACT_LRELU_BWD) {
        const float4 a = reinterpret_cast<const float4*>(pre)[i];
        s.x *= a.x > 0.f ? 1.f : 0.2f; s.y *= a.y > 0.f ? 1.f : 0.2f; s.z *= a.z > 0.f ? 1.f : 0.2f; s.w *= a.w > 0.f ? 1.f : 0.2f;
    } else if (pre) reinterpret_cast<float4*>(pre)[i] = s;
    if (act == ACT_LRELU) {
        s.x = s.x > 0.f ? s.x : 0.2f * s.x; s.y = s.y > 0.f ? s.y : 0.2f * s.y;
        s.z = s.z > 0.f ? s.z : 0.2f * s.z; s.w = s.w > 0.f ? s.w : 0.2f * s.w;
    }
    if (mask) {
        const float4 m = reinterpret_cast<const float4*>(mask)[i];
        s.x *= m.x * mask_scale; s.y *= m.y * mask_scale; s.z *= m.z * mask_scale; s.w *= m.w * mask_scale;
    }
    reinterpret_cast<float4*>(y)[i] = s;
}

// ------------------------------------------------------------------------------------------------ filter gradient
// dW[tap][ci][co] += sum_pos X[pos (+) tap][ci] * DY[pos][co]: the contraction index (positions) is the SLOW index of the
// channels-last tensors, so both operands are staged MN-major (tcgen05 reads either major for tf32; the only MN-major tf32 layout
// is SWIZZLE_128B_BASE32B: 512-byte atoms of 4 positions x 32 channels, the four 32-byte chunks of a 128-byte row XOR-ed with the
// row index): a position's channels are one contiguous 16-byte-vector copy (coalesced LDG.128 -> STS.128, no transposition).
// Tile = [8 k-atoms][channel atoms][4 positions][128 B], LBO = atom stride along channels, SBO = along positions; K = 8 per MMA
// = two k-atoms.
// NVB = N / 32 (channel atoms of the B tile).  KSMALL (critic's first conv: rows = (tap, channel), gathered per element) keeps a
// K-major A tile (row = (tap, channel), 32 positions per 128-byte row).
__device__ __forceinline__ uint64_t make_sdesc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                        // SWIZZLE_128B_BASE32B
    return d;
}

template <int NVB, bool KSMALL>
__global__ void __launch_bounds__(TCG_THREADS, NVB <= 4 ? 2 : 1) tcg_filtergrad_kernel(const __grid_constant__ TcgFilterArgs p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], acc_bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int N = NVB * 32;
    const int nst = p.nstages;
    constexpr uint32_t stage_bytes = 16384u + (uint32_t)N * 128u;
    const TcgFTap tp = p.taps[blockIdx.x];
    const int mt = (int)blockIdx.y % p.Mt, nt = (int)blockIdx.y / p.Mt;
    const int m0 = mt * 128, n0 = nt * N;
    const int nkb_all = (p.rows + 31) >> 5;
    const int kb_lo = (int)((long long)blockIdx.z * nkb_all / p.ksplit);
    const int nkb = (int)((long long)(blockIdx.z + 1) * nkb_all / p.ksplit) - kb_lo;

    if (tid == 0) {
        for (int s = 0; s < nst; ++s) { mbar_init(&full_bar[s], 256); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc(&tmem_slot, tmem_cols_for(N));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 8) {
        // A: MN-major [4 k-atoms][4 channel atoms] (KSMALL: K-major, k step = 32 bytes inside the row); B: MN-major [4][NVB]
        const uint32_t idesc = idesc_tf32(N) | (KSMALL ? 0u : (1u << 15)) | (1u << 16);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % nst;
            mbar_wait(&full_bar[s], (uint32_t)(i / nst) & 1u);
            fence_proxy_async_smem();      // the landed cp.async (generic-proxy) writes -> async proxy
            tc_fence_after();
            if (elect_one()) {
                const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = KSMALL ? make_sdesc(base) + 2 * k : make_sdesc_mn(base + k * 4096, 512, 2048);
                    const uint64_t bd = make_sdesc_mn(base + 16384 + k * (NVB * 1024), 512, NVB * 512);
                    tc_mma_tf32(tmem, ad, bd, idesc, (i | k) ? 1u : 0u);
                }
                tc_commit(&empty_bar[s]);
                if (i == nkb - 1) tc_commit(&acc_bar);
            }
            __syncwarp();
        }
    } else {
        // one position: element offsets of its x / dy rows (channel 0 of this CTA's tiles), or -1
        auto locate = [&](int pos, long long& xoff, long long& yoff) {
            xoff = -1; yoff = -1;
            if (pos >= p.rows) return;
            const int q1 = fast_div(pos, p.mW), w2 = pos - q1 * p.Wc;
            const int q2 = fast_div(q1, p.mH), h2 = q1 - q2 * p.Hc;
            const int b = fast_div(q2, p.mT), t2 = q2 - b * p.Tc;
            const int yt = p.ay * t2 + tp.yt, yh = p.ay * h2 + tp.yh, yw = p.ay * w2 + tp.yw;
            const bool yok = (unsigned)yt < (unsigned)p.Ty && (unsigned)yh < (unsigned)p.Hy && (unsigned)yw < (unsigned)p.Wy;
            if (KSMALL) {          // per-tap validity is resolved per element; xoff packs the tap-free base and the coordinates
                xoff = ((((long long)b * p.Tx + p.ax * t2) * p.Hx + p.ax * h2) * p.Wx + p.ax * w2) * p.Ci;
                xoff = (xoff << 30) | ((long long)(p.ax * t2) << 20) | ((long long)(p.ax * h2) << 10) | (long long)(p.ax * w2);
                if (yok) yoff = ((((long long)b * p.Ty + yt) * p.Hy + yh) * p.Wy + yw) * p.Co + n0;
                return;
            }
            const int xt = p.ax * t2 + tp.xt, xh = p.ax * h2 + tp.xh, xw = p.ax * w2 + tp.xw;
            const bool xok = (unsigned)xt < (unsigned)p.Tx && (unsigned)xh < (unsigned)p.Hx && (unsigned)xw < (unsigned)p.Wx;
            if (xok && yok) {
                xoff = ((((long long)b * p.Tx + xt) * p.Hx + xh) * p.Wx + xw) * p.Ci + m0;
                yoff = ((((long long)b * p.Ty + yt) * p.Hy + yh) * p.Wy + yw) * p.Co + n0;
            }
        };
        // work items = (position, 16-byte channel chunk); a thread's items are tid, tid + 256, ...
        constexpr int CPB = 8 * NVB;                         // chunks per position of the B tile
        constexpr int PPB = 256 / CPB;                       // positions per pass over the 256 producer threads (NVB <= 8: >= 4)
        const int chB = tid % CPB, pB = tid / CPB;
        const int cpa = p.cpa;                               // chunks per position of the A tile (8, 16 or 32)
        const int chA = tid & (cpa - 1), pA = tid / cpa, ppa = 256 / cpa, nA = cpa >> 3;    // nA = items per thread
        // KSMALL: A is K-major, warp w owns rows [16 w, +16), lane = position (2 per lane-pair as before)
        const int psub = lane >> 1, cq = lane & 1;
        const bool a_on = KSMALL ? 16 * warp < p.ksmall : true;

        // MN-major destination of (position k in 0..31, 16-byte channel chunk ch) inside a tile of `atoms` channel atoms
        auto mn_off = [](int atoms, int k, int ch) -> uint32_t {
            const uint32_t r = (uint32_t)k & 3u;
            return (uint32_t)(k >> 2) * (uint32_t)(atoms * 512) + (uint32_t)(ch >> 3) * 512u + r * 128u + (((((uint32_t)ch >> 1) & 3u) ^ r) << 5) +
                   ((uint32_t)ch & 1u) * 16u;
        };
        // asynchronous producers: cp.async straight into the tiles (FP32 bit patterns; the tensor core truncates to tf32), the
        // copies' completion arrives on the stage's full barrier; threads only wait for a free stage
        for (int i = 0; i < nkb; ++i) {
            const int s = i % nst;
            const int pos0 = (kb_lo + i) * 32;
            if (i >= nst) mbar_wait(&empty_bar[s], ((uint32_t)(i / nst) & 1u) ^ 1u);
            const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
            if (!KSMALL) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (j < nA) {
                        long long xo, yo;
                        locate(pos0 + pA + ppa * j, xo, yo);
                        cp_async16(sa + mn_off(4, pA + ppa * j, chA), xo >= 0 ? p.x + xo + 4 * chA : p.x, xo >= 0 ? 16u : 0u);
                    }
            } else if (a_on) {
                long long xo[2], yo[2];
                locate(pos0 + psub, xo[0], yo[0]);
                locate(pos0 + 16 + psub, xo[1], yo[1]);
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int g = 0; g < 2; ++g)
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int m = 16 * warp + 8 * g + 4 * cq + u;      // row = tap * Ci + channel
                            const float* src = p.x;
                            uint32_t nbytes = 0;
                            if (m < p.ksmall && xo[h] >= 0 && yo[h] >= 0) {
                                const int tap = m / p.Ci, c = m - tap * p.Ci;
                                const TcgFTap tq = p.taps[tap];
                                const int xt = (int)((xo[h] >> 20) & 1023) + tq.xt, xh = (int)((xo[h] >> 10) & 1023) + tq.xh, xw = (int)(xo[h] & 1023) + tq.xw;
                                if ((unsigned)xt < (unsigned)p.Tx && (unsigned)xh < (unsigned)p.Hx && (unsigned)xw < (unsigned)p.Wx) {
                                    src = p.x + (xo[h] >> 30) + ((long long)(tq.xt * p.Hx + tq.xh) * p.Wx + tq.xw) * p.Ci + c;
                                    nbytes = 4;
                                }
                            }
                            const uint32_t row = (uint32_t)m, k = (uint32_t)(16 * h + psub);
                            cp_async4(sa + (row >> 3) * 1024u + (row & 7u) * 128u + (((k >> 2) ^ (row & 7u)) << 4) + (k & 3u) * 4u, src, nbytes);
                        }
            }
#pragma unroll
            for (int j = 0; j < NVB; ++j) {
                long long xo, yo;
                locate(pos0 + pB + PPB * j, xo, yo);
                const bool ok = yo >= 0 && (KSMALL || xo >= 0);
                cp_async16(sa + 16384 + mn_off(NVB, pB + PPB * j, chB), ok ? p.dy + yo + 4 * chB : p.dy, ok ? 16u : 0u);
            }
            cp_async_mbar_arrive_noinc(&full_bar[s]);
        }
        // ---- epilogue: accumulator row = input channel (or (tap, channel)), columns = output channels -> dw += D
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
        const int q = warp & 3, half = warp >> 2;
        constexpr int cph = N >= 64 ? N / 2 : 32;
        const int c_lo = half * cph, c_hi = min(N, c_lo + cph);
        const int row = m0 + q * 32 + lane;
        const int row_lim = KSMALL ? p.ksmall : p.Ci;
        float* dst = p.dw + (long long)tp.widx * p.wblk + (long long)row * p.Co + n0;
        for (int c = c_lo; c < c_hi; c += 32) {
            uint32_t v[32];
            tc_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            if (row >= row_lim) continue;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                red_add_v4(dst + c + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem, tmem_cols_for(N)); }
}

// ------------------------------------------------------------------------------------------------ critic first conv, sample-resident
// Scoring mode at nd = 16 (B = tens of thousands): Conv3D(64, 3^3, strides 2, 'valid') on [sample | cond] (Ci = 1 + ncond <= 4).
// A persistent CTA takes one sample at a time: its 24x16x16 field (24 KB) and condition land in shared memory with 16-byte cp.async,
// the im2col A tiles (128 output positions x Kpad, k = tap * Ci + channel) are built from shared memory (no global gathers), the
// packed kernel [64][Kpad] stays resident as the B operand, 5 M tiles per sample (539 positions) with double-buffered A tiles and
// accumulators: workers build tile m + 1 while the tensor core runs tile m, then drain tile m (bias, LeakyReLU, 16-bit store).
struct CriticD1Args {
    const float* sample; const float* cond; const float* wTp; const float* bias; void* out16;
    int B, ncond, kchunks, half_kind;
};
constexpr int D1_ND = 16, D1_T = 24, D1_TO = 11, D1_HO = 7, D1_ROWS = D1_TO * D1_HO * D1_HO;   // 539 output positions
constexpr int D1_MT = (D1_ROWS + 127) / 128;                                                    // 5 M tiles

__global__ void __launch_bounds__(TCG_THREADS, 2) critic_d1_resident_kernel(const __grid_constant__ CriticD1Args p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t a_full[2], a_empty[2], acc_full[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kch = p.kchunks;                                      // = input channels (one k-block per channel)
    const uint32_t a_bytes = (uint32_t)kch * 16384u;                // one A buffer: kch k-blocks of [128 rows x 128 B]
    uint8_t* sB = smem;                                             // [kch][64 rows x 128 B]
    uint8_t* sA = sB + (size_t)kch * 8192;                          // two A buffers
    float* s_in = reinterpret_cast<float*>(sA + 2 * (size_t)a_bytes);   // [24][16][16]
    float* s_cd = s_in + D1_T * D1_ND * D1_ND;                      // [16][16][ncond]
    float* s_bias = s_cd + D1_ND * D1_ND * 3;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 8); mbar_init(&a_empty[i], 1); mbar_init(&acc_full[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) tmem_alloc(&tmem_slot, 128);
    if (tid < 64) s_bias[tid] = p.bias[tid];
    // resident B operand: packed kernel rows (co) x Kpad, K-major swizzled
    for (int i = tid; i < kch * 64 * 8 && tid < 256; i += 256) {
        const int kc = i / 512, r = (i >> 3) & 63, ch = i & 7;
        cp_async16(smem_u32(sB + (size_t)kc * 8192 + (r >> 3) * 1024 + (r & 7) * 128 + ((ch ^ (r & 7)) << 4)), p.wTp + (size_t)r * (kch * 32) + kc * 32 + ch * 4, 16u);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = idesc_tf32(64);
    const int n_mine = ((int)p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;     // samples of this CTA

    if (warp == 8) {
        // tile sequence number g = local sample * 5 + m; buffer = g & 1
        for (int g = 0; g < n_mine * D1_MT; ++g) {
            const int bsel = g & 1;
            mbar_wait(&a_full[bsel], (uint32_t)(g >> 1) & 1u);
            tc_fence_after();
            if (elect_one()) {
                for (int kc = 0; kc < kch; ++kc) {
                    const uint64_t ad = make_sdesc(smem_u32(sA + (size_t)bsel * a_bytes + (size_t)kc * 16384));
                    const uint64_t bd = make_sdesc(smem_u32(sB + (size_t)kc * 8192));
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc_mma_tf32(tmem + bsel * 64, ad + 2 * k, bd + 2 * k, idesc, (kc | k) ? 1u : 0u);
                }
                tc_commit(&a_empty[bsel]);
                tc_commit(&acc_full[bsel]);
            }
            __syncwarp();
        }
    } else {
        // K is channel-major here: k-block c holds the 27 taps of input channel c (k = c * 32 + tap, taps 27..31 zero), so a thread
        // owns ONE output row and one channel: 27 (sample) or 9 (condition, constant over kt) shared-memory reads at stride-2
        // positions (2-way conflicts at most), then eight conflict-free 16-byte stores of its 128-byte tile row
        const int brow = tid & 127, bgrp = tid >> 7;
        auto build = [&](int g, int m) {          // A tile of M tile m into buffer g & 1
            const int bsel = g & 1;
            if (g >= 2) mbar_wait(&a_empty[bsel], ((uint32_t)(g >> 1) & 1u) ^ 1u);
            const int row = m * 128 + brow;
            const bool rv = row < D1_ROWS;
            const int to = row / 49, rem = row - to * 49, ho = rem / 7, wo = rem - ho * 7;
            const float* base = s_in + ((2 * to) * D1_ND + 2 * ho) * D1_ND + 2 * wo;
            const float* cbase = s_cd + ((2 * ho) * D1_ND + 2 * wo) * p.ncond;
            for (int kc = bgrp; kc < kch; kc += 2) {
                float v[32];
#pragma unroll
                for (int t = 0; t < 32; ++t) v[t] = 0.f;
                if (rv) {
                    if (kc == 0) {
#pragma unroll
                        for (int t = 0; t < 27; ++t) v[t] = base[((t / 9) * D1_ND + (t / 3) % 3) * D1_ND + t % 3];
                    } else {
                        float cv[9];
#pragma unroll
                        for (int t = 0; t < 9; ++t) cv[t] = cbase[((t / 3) * D1_ND + t % 3) * p.ncond + (kc - 1)];
#pragma unroll
                        for (int t = 0; t < 27; ++t) v[t] = cv[t % 9];
                    }
                }
                uint8_t* d = sA + (size_t)bsel * a_bytes + (size_t)kc * 16384 + (brow >> 3) * 1024 + (brow & 7) * 128;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    *reinterpret_cast<float4*>(d + ((ch ^ (brow & 7)) << 4)) = make_float4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[bsel]);
        };
        auto drain = [&](int g, int m, long long s) {     // accumulator of tile g -> bias, LeakyReLU, 16-bit store
            const int bsel = g & 1;
            mbar_wait(&acc_full[bsel], (uint32_t)(g >> 1) & 1u);
            tc_fence_after();
            const int q = warp & 3, half = warp >> 2;
            const int row = m * 128 + q * 32 + lane;
            uint32_t v[32];
            tc_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(bsel * 64 + half * 32), v);
            tc_fence_before();
            if (row < D1_ROWS) {
                uint16_t* o = reinterpret_cast<uint16_t*>(p.out16) + ((size_t)s * D1_ROWS + row) * 64 + half * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float x[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) { x[e] = __uint_as_float(v[j + e]) + s_bias[half * 32 + j + e]; x[e] = x[e] > 0.f ? x[e] : 0.2f * x[e]; }
                    uint4 h;
                    if (p.half_kind == RDG_HALF_BF16) {
                        h.x = HalfOps<__nv_bfloat16>::pack(x[0], x[1]); h.y = HalfOps<__nv_bfloat16>::pack(x[2], x[3]);
                        h.z = HalfOps<__nv_bfloat16>::pack(x[4], x[5]); h.w = HalfOps<__nv_bfloat16>::pack(x[6], x[7]);
                    } else {
                        h.x = HalfOps<__half>::pack(x[0], x[1]); h.y = HalfOps<__half>::pack(x[2], x[3]);
                        h.z = HalfOps<__half>::pack(x[4], x[5]); h.w = HalfOps<__half>::pack(x[6], x[7]);
                    }
                    *reinterpret_cast<uint4*>(o + j) = h;
                }
            }
        };
        int g = 0;
        for (long long s = blockIdx.x; s < p.B; s += gridDim.x) {
            // the previous sample's tiles have all been BUILT (not necessarily drained): the input buffer is free
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float* src = p.sample + (size_t)s * D1_T * D1_ND * D1_ND;
            for (int i = tid; i < D1_T * D1_ND * D1_ND / 4; i += 256) cp_async16(smem_u32(s_in + 4 * i), src + 4 * i, 16u);
            for (int i = tid; i < D1_ND * D1_ND * p.ncond; i += 256) cp_async4(smem_u32(s_cd + i), p.cond + (size_t)s * D1_ND * D1_ND * p.ncond + i, 4u);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const int g0 = g;
            build(g0, 0);
            for (int m = 1; m < D1_MT; ++m) {
                build(g0 + m, m);
                drain(g0 + m - 1, m - 1, s);
            }
            drain(g0 + D1_MT - 1, D1_MT - 1, s);
            g += D1_MT;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

__global__ void transpose_blocks_kernel(const float* __restrict__ src, float* __restrict__ dst, int R, int C) {
    __shared__ float tile[32][33];
    const float* s = src + (size_t)blockIdx.z * R * C;
    float* d = dst + (size_t)blockIdx.z * R * C;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < R && c < C) tile[j][threadIdx.x] = s[(size_t)r * C + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < R && c < C) d[(size_t)c * R + r] = tile[threadIdx.x][j];
    }
}

// several transposes in one launch (the derived weight images of one network are rebuilt after every optimizer step, on the
// critical path between two steps: one launch instead of one per layer)
struct TransposeBatch {
    const float* src[4]; float* dst[4];
    int R[4], C[4], bx[4], by[4];        // rows, columns, 32x32 tiles per matrix along C / R
    int begin[5];                        // first linear block of segment i (begin[n] = total); each segment = nblk matrices
    int n;
};
__global__ void transpose_batch_kernel(const __grid_constant__ TransposeBatch tb) {
    __shared__ float tile[32][33];
    int seg = 0;
    while (seg + 1 < tb.n && (int)blockIdx.x >= tb.begin[seg + 1]) ++seg;
    const int R = tb.R[seg], C = tb.C[seg];
    int b = blockIdx.x - tb.begin[seg];
    const int tx = b % tb.bx[seg]; b /= tb.bx[seg];
    const int ty = b % tb.by[seg]; const int z = b / tb.by[seg];
    const float* s = tb.src[seg] + (size_t)z * R * C;
    float* d = tb.dst[seg] + (size_t)z * R * C;
    const int c0 = tx * 32, r0 = ty * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = r0 + j, c = c0 + threadIdx.x;
        if (r < R && c < C) tile[j][threadIdx.x] = s[(size_t)r * C + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, r = r0 + threadIdx.x;
        if (r < R && c < C) d[(size_t)c * R + r] = tile[threadIdx.x][j];
    }
}

// ------------------------------------------------------------------------------------------------ host side
// shared-memory ring depth: two CTAs per SM (about 100 KB each) where the kernel's launch bounds allow it, else one
int pick_stages(int N, bool split) {
    const int stage = (16384 + N * 128) * (split ? 2 : 1);
    const bool two = (N / 32) * (split ? 2 : 1) <= 4;
    const int n = ((two ? 104 : 208) * 1024) / stage;
    return std::max(2, std::min(n, 6));
}
size_t smem_bytes(int N, int nst, bool split) { return (size_t)nst * (16384 + N * 128) * (split ? 2 : 1) + 1024; }

int tile_n_for(int Nt, int mtiles) {
    // The kernels are bound by the L2 -> shared-memory operand stream: bytes per CTA and k-block = 16 KB (A) + 128 N (B), so the
    // widest accumulator minimises the traffic (A is gathered once per N tile); split-K supplies the parallelism instead.
    (void)mtiles;
    for (int N : {256, 128, 64, 32})
        if (Nt % N == 0) return N;
    return 0;
}

// K slices per output tile: minimise (waves of CTAs) x (k-blocks per slice): the launches are a few waves long, so a slice count
// that spills two CTAs into an extra wave costs a third of the kernel.  `slots` = CTAs resident on the whole GPU; every slice
// costs a fixed prologue / epilogue (about four k-blocks, tuned on the training iteration) and one more partial tile to reduce.
int pick_slices(int ctas, int nkb, int slots, int min_kb = 4) {
    int best = 1;
    double best_cost = 1e30;
    const int nmax = std::max(1, nkb / min_kb);
    for (int n = 1; n <= nmax; ++n) {
        const long long waves = ((long long)ctas * n + slots - 1) / slots;
        const double cost = (double)waves * ((nkb + n - 1) / n + 4.0) + 0.05 * n;
        if (cost < best_cost - 1e-9) { best_cost = cost; best = n; }
    }
    return best;
}

template <int NB, bool SPLIT>
int launch_rowgemm_t(const TcgRowArgs& a, dim3 grid, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        RDG_CUDA(cudaFuncSetAttribute(tcg_rowgemm_kernel<NB, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
        attr_done = true;
    }
    tcg_rowgemm_kernel<NB, SPLIT><<<grid, TCG_THREADS, smem_bytes(NB * 32, a.nstages, SPLIT), st>>>(a);
    RDG_LAUNCH_CHECK();
    return 0;
}
int launch_rowgemm_n(const TcgRowArgs& a, dim3 grid, bool split, cudaStream_t st) {
    switch (a.N) {
        case 32:  return split ? launch_rowgemm_t<1, true>(a, grid, st) : launch_rowgemm_t<1, false>(a, grid, st);
        case 64:  return split ? launch_rowgemm_t<2, true>(a, grid, st) : launch_rowgemm_t<2, false>(a, grid, st);
        case 128: return split ? launch_rowgemm_t<4, true>(a, grid, st) : launch_rowgemm_t<4, false>(a, grid, st);
        case 256: return split ? launch_rowgemm_t<8, true>(a, grid, st) : launch_rowgemm_t<8, false>(a, grid, st);
    }
    rdg_set_error("tcg row GEMM: the output channel count must be a multiple of 32");
    return RDG_TCG_E_SHAPE;
}

thread_local TcgArena* t_arena = nullptr;

// partial-slice buffer of `bytes` for a split launch on `st`: the phase's arena when it can serve it, else null (caller allocates)
float* arena_take(size_t bytes, cudaStream_t st) {
    TcgArena* ar = t_arena;
    if (!ar || (ar->bound && ar->stream != st)) return nullptr;
    if (ar->bytes < bytes) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return nullptr;   // no growth while capturing
        if (ar->buf) { if (cudaDeviceSynchronize() != cudaSuccess) return nullptr; cudaFree(ar->buf); ar->buf = nullptr; ar->bytes = 0; }
        const size_t want = bytes + bytes / 4;
        if (cudaMalloc(&ar->buf, want) != cudaSuccess) { cudaGetLastError(); ar->buf = nullptr; return nullptr; }
        ar->bytes = want;
    }
    ar->bound = true; ar->stream = st;
    return ar->buf;
}

int launch_rowgemm(TcgRowArgs& a, int mtiles, int max_kb, cudaStream_t st, const float* bias, float* y, int act, const float* mask,
                   float mask_scale, float* pre, bool split = false) {
    if (split && a.N > 128) a.N = 128;           // the 3xTF32 stage is twice as large
    a.nstages = pick_stages(a.N, split);
    const int ctas = mtiles * (a.Nt / a.N);
    const int slots = (a.N / 32) * (split ? 2 : 1) <= 4 ? 296 : 148;      // resident CTAs: two per SM where the kernel allows it
    const int nslice = max_kb >= 8 ? pick_slices(ctas, max_kb, slots) : 1;
    a.nslice = nslice;
    dim3 grid(mtiles, a.Nt / a.N, nslice);
    if (nslice > 1) {
        const size_t part_bytes = (size_t)nslice * a.out_elems * sizeof(float);
        float* part = arena_take(part_bytes, st);
        const bool own = part == nullptr;
        if (own) RDG_CUDA(cudaMallocAsync(&part, part_bytes, st));
        a.part = part; a.bias = nullptr; a.pre = nullptr; a.mask = nullptr; a.act = ACT_NONE;
        int r = launch_rowgemm_n(a, grid, split, st);
        if (r) return r;
        const long long mn4 = a.out_elems / 4;
        if (ceil_div(mn4, 256) >= 592 || nslice < 4)
            tcg_splitk_epilogue_kernel<1><<<ceil_div(mn4, 256), 256, 0, st>>>(part, nslice, mn4, a.Nt, bias, y, act, mask, mask_scale, pre);
        else if (nslice < 16)
            tcg_splitk_epilogue_kernel<4><<<ceil_div(mn4, 64), 256, 0, st>>>(part, nslice, mn4, a.Nt, bias, y, act, mask, mask_scale, pre);
        else
            tcg_splitk_epilogue_kernel<8><<<ceil_div(mn4, 32), 256, 0, st>>>(part, nslice, mn4, a.Nt, bias, y, act, mask, mask_scale, pre);
        RDG_LAUNCH_CHECK();
        if (own) RDG_CUDA(cudaFreeAsync(part, st));
        return 0;
    }
    a.part = nullptr; a.bias = bias; a.pre = pre; a.mask = mask; a.mask_scale = mask_scale; a.act = act;
    return launch_rowgemm_n(a, grid, split, st);
}

bool fits_i32(long long v) { return v < (1ll << 31); }

}  // namespace

TcgArenaScope::TcgArenaScope(TcgArena* a) : prev(t_arena) {
    if (a) a->bound = false;
    t_arena = a;
}
TcgArenaScope::~TcgArenaScope() { t_arena = prev; }

int tcg_transpose_blocks(const float* src, float* dst, int nblk, int R, int C, cudaStream_t st) {
    dim3 grid(ceil_div(C, 32), ceil_div(R, 32), nblk);
    transpose_blocks_kernel<<<grid, dim3(32, 8), 0, st>>>(src, dst, R, C);
    RDG_LAUNCH_CHECK();
    return 0;
}

int tcg_transpose_blocks_batch(int n, const float* const* src, float* const* dst, const int* nblk, const int* R, const int* C, cudaStream_t st) {
    if (n < 1 || n > 4) { rdg_set_error("tcg_transpose_blocks_batch: 1..4 segments"); return RDG_TCG_E_SHAPE; }
    TransposeBatch tb{};
    tb.n = n;
    int total = 0;
    for (int i = 0; i < n; ++i) {
        tb.src[i] = src[i]; tb.dst[i] = dst[i]; tb.R[i] = R[i]; tb.C[i] = C[i];
        tb.bx[i] = ceil_div(C[i], 32); tb.by[i] = ceil_div(R[i], 32);
        tb.begin[i] = total;
        total += tb.bx[i] * tb.by[i] * nblk[i];
    }
    tb.begin[n] = total;
    transpose_batch_kernel<<<total, dim3(32, 8), 0, st>>>(tb);
    RDG_LAUNCH_CHECK();
    return 0;
}

int tcg_conv_fwd(const float* x, const float* wT, const float* bias, float* y, const ConvGeom& g, int act, const float* mask,
                 float mask_scale, cudaStream_t st, float* pre, int precise) {
    if (g.up || (g.Ci & 3) || (g.Co & 31) || g.KT * g.KH * g.KW > 64) { rdg_set_error("tcg_conv_fwd: unsupported geometry"); return RDG_TCG_E_SHAPE; }
    const long long rows = (long long)g.B * g.To * g.Ho * g.Wo;
    if (rows == 0) return 0;
    if (!fits_i32((long long)g.B * g.Ti * g.Hi * g.Wi * g.Ci) || !fits_i32(rows * g.Co)) { rdg_set_error("tcg_conv_fwd: tensor too large"); return RDG_TCG_E_SHAPE; }
    TcgRowArgs a{};
    a.src = x; a.w = wT; a.out = y;
    a.a = g.stride; a.os = 1;
    a.Ts = g.Ti; a.Hs = g.Hi; a.Ws = g.Wi; a.Cs = g.Ci;
    a.To = g.To; a.Ho = g.Ho; a.Wo = g.Wo; a.Nt = g.Co;
    a.Kc = g.Ci; a.kchunks = ceil_div(g.Ci, 32); a.wrow = g.Ci; a.wtap = (long long)g.Co * g.Ci;
    a.out_elems = rows * g.Co;
    const int mtiles = ceil_div(rows, 128);
    a.N = tile_n_for(g.Co, mtiles);
    a.nclass = 1;
    a.cls[0] = TcgClass{g.To, g.Ho, g.Wo, (int)rows, 0, 0, 0, 0, 0, g.KT * g.KH * g.KW};
    int n = 0;
    for (int kt = 0; kt < g.KT; ++kt)
        for (int kh = 0; kh < g.KH; ++kh)
            for (int kw = 0; kw < g.KW; ++kw, ++n) a.taps[n] = TcgTap{(int8_t)(kt - g.pt), (int8_t)(kh - g.ph), (int8_t)(kw - g.pw), 0, n};
    return launch_rowgemm(a, mtiles, n * a.kchunks, st, bias, y, act, mask, mask_scale, pre, precise != 0);
}

// ---- few-channel input (the critic's first conv: Ci = 1 + ncond): K = taps * Ci elements gathered one by one
int tcg_smallci_kpad(int taps, int Ci) { return (taps * Ci + 31) / 32 * 32; }

namespace {
__global__ void pack_smallci_kernel(const float* __restrict__ w, float* __restrict__ wTp, int K, int Kpad, int Co) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Co * Kpad) return;
    const int n = i / Kpad, k = i - n * Kpad;
    wTp[i] = k < K ? w[(size_t)k * Co + n] : 0.f;
}
}  // namespace

namespace {
__global__ void pack_smallci_chmajor_kernel(const float* __restrict__ w, float* __restrict__ wTq, int taps, int Ci, int Co) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Co * Ci * 32) return;
    const int n = i / (Ci * 32), k = i - n * (Ci * 32), c = k >> 5, tap = k & 31;
    wTq[i] = tap < taps ? w[((size_t)tap * Ci + c) * Co + n] : 0.f;
}
}  // namespace

// channel-major variant for the sample-resident kernel: [Co][Ci * 32], k = channel * 32 + tap (taps <= 32)
int tcg_pack_smallci_weights_chmajor(const float* w, float* wTq, int taps, int Ci, int Co, cudaStream_t st) {
    pack_smallci_chmajor_kernel<<<ceil_div((long long)Co * Ci * 32, 256), 256, 0, st>>>(w, wTq, taps, Ci, Co);
    RDG_LAUNCH_CHECK();
    return 0;
}

int tcg_pack_smallci_weights(const float* w, float* wTp, int taps, int Ci, int Co, cudaStream_t st) {
    const int Kpad = tcg_smallci_kpad(taps, Ci);
    pack_smallci_kernel<<<ceil_div((long long)Co * Kpad, 256), 256, 0, st>>>(w, wTp, taps * Ci, Kpad, Co);
    RDG_LAUNCH_CHECK();
    return 0;
}

int tcg_conv_fwd_smallci(const float* x, const float* wTp, const float* bias, float* y, const ConvGeom& g, int act, const float* mask,
                         float mask_scale, cudaStream_t st, float* pre, int precise) {
    const int taps = g.KT * g.KH * g.KW;
    if (g.up || g.Ci > 4 || (g.Co & 31) || taps > 64 || taps * g.Ci > 128) { rdg_set_error("tcg_conv_fwd_smallci: unsupported geometry"); return RDG_TCG_E_SHAPE; }
    const long long rows = (long long)g.B * g.To * g.Ho * g.Wo;
    if (rows == 0) return 0;
    if (!fits_i32((long long)g.B * g.Ti * g.Hi * g.Wi * g.Ci) || !fits_i32(rows * g.Co)) { rdg_set_error("tcg_conv_fwd_smallci: tensor too large"); return RDG_TCG_E_SHAPE; }
    TcgRowArgs a{};
    a.src = x; a.w = wTp; a.out = y;
    a.a = g.stride; a.os = 1;
    a.Ts = g.Ti; a.Hs = g.Hi; a.Ws = g.Wi; a.Cs = g.Ci;
    a.To = g.To; a.Ho = g.Ho; a.Wo = g.Wo; a.Nt = g.Co;
    a.ksmall = taps * g.Ci;
    a.Kc = tcg_smallci_kpad(taps, g.Ci); a.kchunks = a.Kc / 32; a.wrow = a.Kc; a.wtap = 0;
    a.out_elems = rows * g.Co;
    const int mtiles = ceil_div(rows, 128);
    a.N = tile_n_for(g.Co, mtiles);
    a.nclass = 1;
    a.cls[0] = TcgClass{g.To, g.Ho, g.Wo, (int)rows, 0, 0, 0, 0, 0, taps};
    int n = 0;
    for (int kt = 0; kt < g.KT; ++kt)
        for (int kh = 0; kh < g.KH; ++kh)
            for (int kw = 0; kw < g.KW; ++kw, ++n) a.taps[n] = TcgTap{(int8_t)(kt - g.pt), (int8_t)(kh - g.ph), (int8_t)(kw - g.pw), 0, 0};
    return launch_rowgemm(a, mtiles, a.kchunks, st, bias, y, act, mask, mask_scale, pre, precise != 0);
}

// The critic's first conv in the 16-bit scoring mode (critic_tc.cu consumes 16-bit activations): sample [B,24,nd,nd] and cond
// [B,nd,nd,ncond] are concatenated inside the gather (:275-282), single-pass tf32, LeakyReLU, 16-bit channels-last output.
int tcg_critic_first_conv16(int half_kind, const float* sample, const float* cond, const float* wTp, const float* wTq, const float* bias,
                            void* out16, const ConvGeom& g, cudaStream_t st) {
    const int taps = g.KT * g.KH * g.KW;
    if (g.up || g.Ci > 4 || g.Ci < 2 || (g.Co & 31) || taps > 64 || taps * g.Ci > 128) { rdg_set_error("tcg_critic_first_conv16: unsupported geometry"); return RDG_TCG_E_SHAPE; }
    const long long rows = (long long)g.B * g.To * g.Ho * g.Wo;
    if (rows == 0) return 0;
    if (!fits_i32((long long)g.B * g.Ti * g.Hi * g.Wi * g.Ci) || !fits_i32(rows * g.Co)) { rdg_set_error("tcg_critic_first_conv16: tensor too large"); return RDG_TCG_E_SHAPE; }

    // nd = 16: sample-resident kernel (shared-memory im2col, resident weights); other domains: the general gather kernel below
    if (wTq && taps == 27 && g.Ti == D1_T && g.Hi == D1_ND && g.Wi == D1_ND && g.Co == 64 && g.stride == 2 && g.pt == 0 && g.ph == 0 && g.pw == 0 && g.To == D1_TO &&
        !(getenv("RDG_CRITIC_D1") && strcmp(getenv("RDG_CRITIC_D1"), "gather") == 0)) {
        CriticD1Args a{};
        a.sample = sample; a.cond = cond; a.wTp = wTq; a.bias = bias; a.out16 = out16;
        a.B = g.B; a.ncond = g.Ci - 1; a.kchunks = g.Ci; a.half_kind = half_kind;
        const size_t smem = 1024 + (size_t)a.kchunks * 8192 + 2 * (size_t)a.kchunks * 16384 + (D1_T * D1_ND * D1_ND + D1_ND * D1_ND * 3 + 64) * 4;
        static bool attr_done = false;
        if (!attr_done) {
            RDG_CUDA(cudaFuncSetAttribute(critic_d1_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
            attr_done = true;
        }
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        const int ctas = std::min(g.B, sms * (smem <= 110 * 1024 ? 2 : 1));
        critic_d1_resident_kernel<<<ctas, TCG_THREADS, smem, st>>>(a);
        RDG_LAUNCH_CHECK();
        return 0;
    }
    TcgRowArgs a{};
    a.src = sample; a.src2 = cond; a.w = wTp; a.out = nullptr; a.out16 = out16; a.half_kind = half_kind;
    a.a = g.stride; a.os = 1;
    a.Ts = g.Ti; a.Hs = g.Hi; a.Ws = g.Wi; a.Cs = g.Ci;
    a.To = g.To; a.Ho = g.Ho; a.Wo = g.Wo; a.Nt = g.Co;
    a.ksmall = taps * g.Ci;
    a.Kc = tcg_smallci_kpad(taps, g.Ci); a.kchunks = a.Kc / 32; a.wrow = a.Kc; a.wtap = 0;
    a.out_elems = rows * g.Co;
    const int mtiles = ceil_div(rows, 128);
    a.N = tile_n_for(g.Co, mtiles);
    a.nclass = 1;
    a.cls[0] = TcgClass{g.To, g.Ho, g.Wo, (int)rows, 0, 0, 0, 0, 0, taps};
    int n = 0;
    for (int kt = 0; kt < g.KT; ++kt)
        for (int kh = 0; kh < g.KH; ++kh)
            for (int kw = 0; kw < g.KW; ++kw, ++n) a.taps[n] = TcgTap{(int8_t)(kt - g.pt), (int8_t)(kh - g.ph), (int8_t)(kw - g.pw), 0, 0};
    a.nstages = pick_stages(a.N, false);
    a.nslice = 1;
    a.bias = bias; a.act = ACT_LRELU; a.mask_scale = 1.f;
    return launch_rowgemm_n(a, dim3(mtiles, a.Nt / a.N, 1), false, st);
}

int tcg_conv_bwd_data(const float* dy, const float* w, float* dx, const ConvGeom& g, cudaStream_t st, const float* pre_in, const float* mask,
                      float mask_scale) {
    if (g.up || (g.Co & 3) || (g.Ci & 31) || (g.stride != 1 && g.stride != 2) || g.KT * g.KH * g.KW > 64) {
        rdg_set_error("tcg_conv_bwd_data: unsupported geometry"); return RDG_TCG_E_SHAPE;
    }
    const long long rows = (long long)g.B * g.Ti * g.Hi * g.Wi;
    if (rows == 0) return 0;
    if (!fits_i32((long long)g.B * g.To * g.Ho * g.Wo * g.Co) || !fits_i32(rows * g.Ci)) { rdg_set_error("tcg_conv_bwd_data: tensor too large"); return RDG_TCG_E_SHAPE; }
    TcgRowArgs a{};
    a.src = dy; a.w = w; a.out = dx;
    a.a = 1; a.os = g.stride;
    a.Ts = g.To; a.Hs = g.Ho; a.Ws = g.Wo; a.Cs = g.Co;
    a.To = g.Ti; a.Ho = g.Hi; a.Wo = g.Wi; a.Nt = g.Ci;
    a.Kc = g.Co; a.kchunks = ceil_div(g.Co, 32); a.wrow = g.Co; a.wtap = (long long)g.Ci * g.Co;
    a.out_elems = rows * g.Ci;
    const int s = g.stride;
    int ncls = 0, ntap = 0, tiles = 0, max_taps = 0;
    for (int qt = 0; qt < s; ++qt)
        for (int qh = 0; qh < s; ++qh)
            for (int qw = 0; qw < s; ++qw) {
                const int Tc = (g.Ti - qt + s - 1) / s, Hc = (g.Hi - qh + s - 1) / s, Wc = (g.Wi - qw + s - 1) / s;
                if (Tc <= 0 || Hc <= 0 || Wc <= 0) continue;
                TcgClass c{Tc, Hc, Wc, g.B * Tc * Hc * Wc, tiles, qt, qh, qw, ntap, 0};
                for (int kt = 0; kt < g.KT; ++kt) {
                    if ((qt + g.pt - kt) % s) continue;
                    for (int kh = 0; kh < g.KH; ++kh) {
                        if ((qh + g.ph - kh) % s) continue;
                        for (int kw = 0; kw < g.KW; ++kw) {
                            if ((qw + g.pw - kw) % s) continue;
                            a.taps[ntap++] = TcgTap{(int8_t)((qt + g.pt - kt) / s), (int8_t)((qh + g.ph - kh) / s), (int8_t)((qw + g.pw - kw) / s), 0,
                                                    (kt * g.KH + kh) * g.KW + kw};
                            ++c.tap_count;
                        }
                    }
                }
                if (c.tap_count == 0) {   // positions no output reads: the gradient is zero; one all-invalid tap keeps the tile well-formed
                    a.taps[ntap++] = TcgTap{(int8_t)-128, (int8_t)-128, (int8_t)-128, 0, 0};
                    c.tap_count = 1;
                }
                max_taps = std::max(max_taps, c.tap_count);
                tiles += ceil_div(c.rows, 128);
                a.cls[ncls++] = c;
            }
    a.nclass = ncls;
    a.N = tile_n_for(g.Ci, tiles);
    return launch_rowgemm(a, tiles, max_taps * a.kchunks, st, nullptr, dx, pre_in ? ACT_LRELU_BWD : ACT_NONE, pre_in ? mask : nullptr, mask_scale,
                          const_cast<float*>(pre_in));
}

int tcg_folded_fwd(const float* x, const float* wfT, const float* bias, float* y, const ConvGeom& g, cudaStream_t st, int precise) {
    if (!g.up || (g.Ci & 31) || (g.Co & 31)) { rdg_set_error("tcg_folded_fwd: unsupported geometry"); return RDG_TCG_E_SHAPE; }
    const long long rows = (long long)g.B * g.Ti * g.Hi * g.Wi;
    if (rows == 0) return 0;
    if (!fits_i32(rows * g.Ci) || !fits_i32(rows * 8 * g.Co)) { rdg_set_error("tcg_folded_fwd: tensor too large"); return RDG_TCG_E_SHAPE; }
    TcgRowArgs a{};
    a.src = x; a.w = wfT; a.out = y;
    a.a = 1; a.os = 2;
    a.Ts = g.Ti; a.Hs = g.Hi; a.Ws = g.Wi; a.Cs = g.Ci;
    a.To = 2 * g.Ti; a.Ho = 2 * g.Hi; a.Wo = 2 * g.Wi; a.Nt = g.Co;
    a.Kc = g.Ci; a.kchunks = g.Ci / 32; a.wrow = g.Ci; a.wtap = (long long)g.Co * g.Ci;
    a.out_elems = rows * 8 * g.Co;
    const int tpc = ceil_div(rows, 128);
    for (int p = 0; p < 8; ++p) {
        const int pt = p >> 2, ph = (p >> 1) & 1, pw = p & 1;
        a.cls[p] = TcgClass{g.Ti, g.Hi, g.Wi, (int)rows, p * tpc, pt, ph, pw, p * 8, 8};
        for (int t = 0; t < 8; ++t) {
            const int at = t >> 2, ah = (t >> 1) & 1, aw = t & 1;
            a.taps[p * 8 + t] = TcgTap{(int8_t)(at - 1 + pt), (int8_t)(ah - 1 + ph), (int8_t)(aw - 1 + pw), 0, p * 8 + t};
        }
    }
    a.nclass = 8;
    a.N = tile_n_for(g.Co, 8 * tpc);
    return launch_rowgemm(a, 8 * tpc, 8 * a.kchunks, st, bias, y, ACT_NONE, nullptr, 1.f, nullptr, precise != 0);
}

int tcg_folded_bwd_data(const float* dy, const float* wf, float* dx, const ConvGeom& g, cudaStream_t st) {
    if (!g.up || (g.Ci & 31) || (g.Co & 31)) { rdg_set_error("tcg_folded_bwd_data: unsupported geometry"); return RDG_TCG_E_SHAPE; }
    const long long rows = (long long)g.B * g.Ti * g.Hi * g.Wi;
    if (rows == 0) return 0;
    if (!fits_i32(rows * g.Ci) || !fits_i32(rows * 8 * g.Co)) { rdg_set_error("tcg_folded_bwd_data: tensor too large"); return RDG_TCG_E_SHAPE; }
    TcgRowArgs a{};
    a.src = dy; a.w = wf; a.out = dx;
    a.a = 2; a.os = 1;
    a.Ts = 2 * g.Ti; a.Hs = 2 * g.Hi; a.Ws = 2 * g.Wi; a.Cs = g.Co;
    a.To = g.Ti; a.Ho = g.Hi; a.Wo = g.Wi; a.Nt = g.Ci;
    a.Kc = g.Co; a.kchunks = g.Co / 32; a.wrow = g.Co; a.wtap = (long long)g.Ci * g.Co;
    a.out_elems = rows * g.Ci;
    const int mtiles = ceil_div(rows, 128);
    a.cls[0] = TcgClass{g.Ti, g.Hi, g.Wi, (int)rows, 0, 0, 0, 0, 0, 64};
    for (int p = 0; p < 8; ++p) {
        const int pt = p >> 2, ph = (p >> 1) & 1, pw = p & 1;
        for (int t = 0; t < 8; ++t) {
            const int at = t >> 2, ah = (t >> 1) & 1, aw = t & 1;
            // x[q] feeds output 2(q - d) + phase with d = a - 1 + phase  ->  source offset phase - 2d on the high-res grid
            a.taps[p * 8 + t] = TcgTap{(int8_t)(pt - 2 * (at - 1 + pt)), (int8_t)(ph - 2 * (ah - 1 + ph)), (int8_t)(pw - 2 * (aw - 1 + pw)), 0, p * 8 + t};
        }
    }
    a.nclass = 1;
    a.N = tile_n_for(g.Ci, mtiles);
    return launch_rowgemm(a, mtiles, 64 * a.kchunks, st, nullptr, dx, ACT_NONE, nullptr, 1.f, nullptr);
}

namespace {
template <int NVB, bool KSMALL>
int launch_filtergrad_t(const TcgFilterArgs& a, dim3 grid, cudaStream_t st) {
    static bool attr_done = false;
    if (!attr_done) {
        RDG_CUDA(cudaFuncSetAttribute(tcg_filtergrad_kernel<NVB, KSMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048));
        attr_done = true;
    }
    tcg_filtergrad_kernel<NVB, KSMALL><<<grid, TCG_THREADS, smem_bytes(NVB * 32, a.nstages, false), st>>>(a);
    RDG_LAUNCH_CHECK();
    return 0;
}
unsigned long long fast_div_mul(int d) { return ((1ull << 40) + (unsigned long long)d - 1) / (unsigned long long)d; }
int launch_filtergrad(TcgFilterArgs& a, int ntap, cudaStream_t st) {
    a.Mt = a.ksmall ? 1 : ceil_div(a.Ci, 128);
    if (a.rows >= (1 << 24) || a.Wc >= 256 || a.Hc >= 256 || a.Tc >= 256) { rdg_set_error("tcg filter gradient: more than 2^24 positions"); return RDG_TCG_E_SHAPE; }
    a.mW = fast_div_mul(a.Wc); a.mH = fast_div_mul(a.Hc); a.mT = fast_div_mul(a.Tc);
    a.cpa = std::min(a.Ci, 128) / 4;
    if (!a.ksmall && ((a.Ci & 31) || (a.cpa & (a.cpa - 1)) || (a.Ci > 128 && a.Ci % 128))) { rdg_set_error("tcg filter gradient: Ci must be 32, 64 or a multiple of 128"); return RDG_TCG_E_SHAPE; }
    int N = 0;
    for (int n : {256, 128, 64, 32})
        if (a.Co % n == 0) { N = n; break; }     // widest B tile: the A tile is fetched once per N tile (operand-stream bound)
    if (!N) { rdg_set_error("tcg filter gradient: Co must be a multiple of 32"); return RDG_TCG_E_SHAPE; }
    a.N = N;
    a.nstages = pick_stages(N, false);
    const int ctas = ntap * a.Mt * (a.Co / N);
    const int nkb = (a.rows + 31) / 32;
    const int ks = pick_slices(ctas, nkb, N <= 128 ? 296 : 148);
    a.ksplit = ks;
    dim3 grid(ntap, a.Mt * (a.Co / N), ks);
    if (a.ksmall) {
        switch (N) {
            case 32:  return launch_filtergrad_t<1, true>(a, grid, st);
            case 64:  return launch_filtergrad_t<2, true>(a, grid, st);
            case 128: return launch_filtergrad_t<4, true>(a, grid, st);
            default:  return launch_filtergrad_t<8, true>(a, grid, st);
        }
    }
    switch (N) {
        case 32:  return launch_filtergrad_t<1, false>(a, grid, st);
        case 64:  return launch_filtergrad_t<2, false>(a, grid, st);
        case 128: return launch_filtergrad_t<4, false>(a, grid, st);
        default:  return launch_filtergrad_t<8, false>(a, grid, st);
    }
}
}  // namespace

int tcg_conv_bwd_filter_smallci(const float* x, const float* dy, float* dw, const ConvGeom& g, cudaStream_t st) {
    const int taps = g.KT * g.KH * g.KW;
    if (g.up || g.Ci > 4 || (g.Co & 31) || taps > 64 || taps * g.Ci > 128) { rdg_set_error("tcg_conv_bwd_filter_smallci: unsupported geometry"); return RDG_TCG_E_SHAPE; }
    const long long rows = (long long)g.B * g.To * g.Ho * g.Wo;
    if (rows == 0) return 0;
    if (!fits_i32(rows) || !fits_i32((long long)g.B * g.Ti * g.Hi * g.Wi * g.Ci)) { rdg_set_error("tcg_conv_bwd_filter_smallci: tensor too large"); return RDG_TCG_E_SHAPE; }
    TcgFilterArgs a{};
    a.x = x; a.dy = dy; a.dw = dw;
    a.Tc = g.To; a.Hc = g.Ho; a.Wc = g.Wo; a.rows = (int)rows;
    a.ax = g.stride; a.Tx = g.Ti; a.Hx = g.Hi; a.Wx = g.Wi; a.Ci = g.Ci;
    a.ay = 1; a.Ty = g.To; a.Hy = g.Ho; a.Wy = g.Wo; a.Co = g.Co;
    a.ksmall = taps * g.Ci;
    a.wblk = 0;
    int n = 0;
    for (int kt = 0; kt < g.KT; ++kt)
        for (int kh = 0; kh < g.KH; ++kh)
            for (int kw = 0; kw < g.KW; ++kw, ++n)
                a.taps[n] = TcgFTap{(int8_t)(kt - g.pt), (int8_t)(kh - g.ph), (int8_t)(kw - g.pw), 0, 0, 0, 0};
    return launch_filtergrad(a, 1, st);
}

int tcg_conv_bwd_filter(const float* x, const float* dy, float* dw, const ConvGeom& g, cudaStream_t st) {
    if (g.up || (g.Ci & 15) || (g.Co & 31) || g.KT * g.KH * g.KW > 64) { rdg_set_error("tcg_conv_bwd_filter: unsupported geometry"); return RDG_TCG_E_SHAPE; }
    const long long rows = (long long)g.B * g.To * g.Ho * g.Wo;
    if (rows == 0) return 0;
    if (!fits_i32(rows)) { rdg_set_error("tcg_conv_bwd_filter: tensor too large"); return RDG_TCG_E_SHAPE; }
    TcgFilterArgs a{};
    a.x = x; a.dy = dy; a.dw = dw;
    a.Tc = g.To; a.Hc = g.Ho; a.Wc = g.Wo; a.rows = (int)rows;
    a.ax = g.stride; a.Tx = g.Ti; a.Hx = g.Hi; a.Wx = g.Wi; a.Ci = g.Ci;
    a.ay = 1; a.Ty = g.To; a.Hy = g.Ho; a.Wy = g.Wo; a.Co = g.Co;
    a.wblk = (long long)g.Ci * g.Co;
    int n = 0;
    for (int kt = 0; kt < g.KT; ++kt)
        for (int kh = 0; kh < g.KH; ++kh)
            for (int kw = 0; kw < g.KW; ++kw, ++n)
                a.taps[n] = TcgFTap{(int8_t)(kt - g.pt), (int8_t)(kh - g.ph), (int8_t)(kw - g.pw), 0, 0, 0, (int16_t)n};
    return launch_filtergrad(a, n, st);
}

int tcg_folded_bwd_filter(const float* x, const float* dy, float* dwf, const ConvGeom& g, cudaStream_t st) {
    if (!g.up || (g.Ci & 15) || (g.Co & 31)) { rdg_set_error("tcg_folded_bwd_filter: unsupported geometry"); return RDG_TCG_E_SHAPE; }
    const long long rows = (long long)g.B * g.Ti * g.Hi * g.Wi;
    if (rows == 0) return 0;
    if (!fits_i32(rows)) { rdg_set_error("tcg_folded_bwd_filter: tensor too large"); return RDG_TCG_E_SHAPE; }
    TcgFilterArgs a{};
    a.x = x; a.dy = dy; a.dw = dwf;
    a.Tc = g.Ti; a.Hc = g.Hi; a.Wc = g.Wi; a.rows = (int)rows;
    a.ax = 1; a.Tx = g.Ti; a.Hx = g.Hi; a.Wx = g.Wi; a.Ci = g.Ci;
    a.ay = 2; a.Ty = 2 * g.Ti; a.Hy = 2 * g.Hi; a.Wy = 2 * g.Wi; a.Co = g.Co;
    a.wblk = (long long)g.Ci * g.Co;
    for (int p = 0; p < 8; ++p) {
        const int pt = p >> 2, ph = (p >> 1) & 1, pw = p & 1;
        for (int t = 0; t < 8; ++t) {
            const int at = t >> 2, ah = (t >> 1) & 1, aw = t & 1;
            a.taps[p * 8 + t] = TcgFTap{(int8_t)(at - 1 + pt), (int8_t)(ah - 1 + ph), (int8_t)(aw - 1 + pw), (int8_t)pt, (int8_t)ph, (int8_t)pw,
                                        (int16_t)(p * 8 + t)};
        }
    }
    return launch_filtergrad(a, 64, st);
}
