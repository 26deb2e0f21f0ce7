// Tensor-core generator layers for sm_100a: UpSampling3D(2) + Conv3D(3^3,'same') + bias +
// PixelNormalization + LeakyReLU(0.2) (gan_train_cwgangp_pixelnorm.py:330-343) as ONE kernel.
//
// Algorithm: the nearest x2 upsample is folded into the weights: each of the 8
// output phases (pt,ph,pw) is a 2x2x2 convolution on the LOW-RES grid.  Per phase and tap the
// contribution is a plain GEMM  D[128 positions, Cout] += A[128, Cin] * Wf[phase][tap][Cin, Cout]
// where A is the input activation tile shifted by the tap offset.  A tiles are fetched with 5-D
// TMA box loads straight from the channels-last activation tensor (zero fill outside the grid
// implements the conv padding), weight tiles with 1-D bulk copies of host-pre-swizzled images;
// both land in 128B-swizzled K-major shared memory and feed tcgen05.mma (kind::f16, M=128,
// N=Cout, K=16) accumulating in TMEM.  NPH phases share one pass so that an A tile is reused by
// every phase that needs that offset.  The epilogue warps read the accumulators back with
// tcgen05.ld, apply bias + PixelNorm + LeakyReLU in FP32 and store 16-bit channels-last output.
//
// Warp roles (224 threads): warp 0 = activation (A) TMA producer, warp 1 = TMEM alloc + MMA issuer,
// warp 2 = weight (B) bulk-copy producer, warps 3..6 = epilogue (one TMEM lane quarter each).
#include "rdg_common.cuh"
#include "gen_tc.h"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <cstdlib>

using namespace rdg_tc;

namespace {

// ------------------------------------------------------------------ step enumeration
// One source of truth for the K-loop order, shared by the producer and the MMA issuer.
// For pass `pass` (phases pass*NPH .. pass*NPH+NPH-1) walk the 27 low-res offsets; an offset
// (dt,dh,dw) serves phase p=(pt,ph,pw) with tap a=(dt+1-pt, dh+1-ph, dw+1-pw) if all in {0,1}.
// SKIP_OOB: offsets whose whole activation tile lies in the zero padding (t+dt outside the grid) are dropped.
// With weight multicast (cluster > 1) every CTA of the cluster must walk the SAME weight sequence, so the step
// is kept (oob = true): no activation load, no MMA, but the weight stage is still consumed.
// The w offsets are walked in the order 0, -1, +1.  MERGE (two w-phases per pass on CTA pairs): the dw = 0 view serves both
// phases of the pass, so its two weight tiles form ONE N = 2*Cout operand; such a step is reported per 64-channel chunk
// as on_b(0, tile of phase 0, oob, k, tile of phase 1) with k >= 0 (k = -1: ordinary single-phase stage of kc chunks).
// Coming first in every (dt, dh) group, the merged step finds both accumulators in the same started / not-started state.
template <int NPH, bool SKIP_OOB, bool MERGE, typename FA, typename FB>
__device__ __forceinline__ void for_each_step(int pass, int t, int T, int nchunk, int kc, FA&& on_a, FB&& on_b) {
    for (int o = 0; o < 27; ++o) {
        const int dt = o / 9 - 1, dh = (o / 3) % 3 - 1, dw = (o % 3 == 0) ? 0 : (o % 3 == 1 ? -1 : 1);
        const bool oob = t + dt < 0 || t + dt >= T;   // whole tile in the zero padding
        if (SKIP_OOB && oob) continue;
        uint32_t mask = 0;
#pragma unroll
        for (int s = 0; s < NPH; ++s) {
            const int p = pass * NPH + s;
            const int at = dt + 1 - (p >> 2), ah = dh + 1 - ((p >> 1) & 1), aw = dw + 1 - (p & 1);
            if ((unsigned)at < 2u && (unsigned)ah < 2u && (unsigned)aw < 2u) mask |= 1u << s;
        }
        if (!mask) continue;
        for (int c = 0; c < nchunk; c += kc) {       // kc consecutive 64-channel chunks per stage
            on_a(dt, dh, dw, c, oob);
            if (MERGE && mask == 3u) {
                const int p0 = pass * NPH, p1 = p0 + 1;
                const int a0 = ((dt + 1 - (p0 >> 2)) << 2) | ((dh + 1 - ((p0 >> 1) & 1)) << 1) | (dw + 1 - (p0 & 1));
                const int a1 = ((dt + 1 - (p1 >> 2)) << 2) | ((dh + 1 - ((p1 >> 1) & 1)) << 1) | (dw + 1 - (p1 & 1));
                for (int k = 0; k < kc; ++k) on_b(0, (p0 * 8 + a0) * nchunk + c + k, oob, k, (p1 * 8 + a1) * nchunk + c + k);
                continue;
            }
#pragma unroll
            for (int s = 0; s < NPH; ++s) {
                if (!(mask & (1u << s))) continue;
                const int p = pass * NPH + s;
                const int a = ((dt + 1 - (p >> 2)) << 2) | ((dh + 1 - ((p >> 1) & 1)) << 1) | (dw + 1 - (p & 1));
                on_b(s, (p * 8 + a) * nchunk + c, oob, -1, 0);
            }
        }
    }
}

constexpr int kThreads = 224;       // warp 0: A producer, 1: MMA issuer, 2: B producer, 3..6: epilogue
constexpr int kATile = 128 * 128;   // 128 rows x 64 x 2 B (one 64-channel chunk)

// KC = 64-channel chunks per pipeline stage: a stage carries K = 64*KC, i.e. 4*KC MMAs per weight tile,
// which amortises the mbarrier round trip of the single-thread producer / issuer loops.
// CG = CTAs sharing every MMA (tcgen05 cta_group): with CG = 2 a CTA's shared memory holds HALF the rows of each weight tile.
template <int COUT, int NPH, int KC, int CG = 1> struct TcCfg {
    static constexpr int kBTile = COUT * 128 / CG;       // this CTA's part of one (phase, tap, chunk) weight tile
    static constexpr int kAStage = KC * kATile;
    static constexpr int kBStage = KC * kBTile;
    static constexpr int kAccCols = NPH * COUT;
    static constexpr int kAccStages = (kAccCols * 2 <= 512) ? 2 : 1;
    static constexpr int kFuseBytes = COUT == 64 ? kATile + 4096 : 0;
    static constexpr int kAStages = KC == 1 ? (CG == 2 ? 6 : 4) : (CG == 2 ? 4 : 3);   // CTA pairs: half-size weight stages leave room for a deeper activation ring
    static constexpr int kBudget = 218 * 1024 - kFuseBytes - kAStages * kAStage;
    static constexpr int kBStages = kBudget / kBStage > 8 * CG ? 8 * CG : kBudget / kBStage;
    static constexpr int kSmem = 1024 + kAStages * kAStage + kBStages * kBStage + kFuseBytes + COUT * 4 + 512;
    static_assert(kAccCols * kAccStages <= 512, "TMEM overflow");
    static_assert(kBStages >= 2, "need at least 2 weight stages");
    static_assert(kSmem <= 227 * 1024, "shared memory overflow");
};

// FUSE (Cout=64 layer only): the epilogue does not store the activations y; it writes them as a
// 128B-swizzled A tile to shared memory and multiplies by the output conv's 27 tap vectors
// (W4 as a [32 x 64] B tile) on the tensor core:  P[pos][tap] = sum_c y[pos][c] * w4[tap][c].
// P (f32, 32 per position) goes to HBM instead of y; gather_softmax_kernel then forms
// logit[q] = b + sum_tap P[q + offset(tap)][tap] and the softmax over the 24 hours.
// CG = 2: CTA pairs (tcgen05 cta_group::2).  The two CTAs of a cluster work on two position tiles that differ only in
// the sample block (same t and h block, so both walk the same step sequence); every MMA is M = 256 (128 rows per CTA),
// issued by the leader (rank 0); each CTA's shared memory holds its own activation tiles and HALF of the rows of every
// weight tile, so the weight-operand reads and weight writes per SM halve.  Full barriers live in the leader (both
// CTAs' TMA loads complete on them: .cta_group::2 + mapa address), releases are multicast commits, the peer's epilogue
// releases accumulators by a remote arrive.  Primitives verified by tools/umma_2cta_probe.cu.
template <typename HT, int COUT, int NPH, int KC, bool FUSE, int CG>
__global__ void __launch_bounds__(kThreads, 1)
tc_upconv_pixelnorm_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_w, TcConvArgs args) {
    using Cfg = TcCfg<COUT, NPH, KC, CG>;
    constexpr bool kSkip = true;
    constexpr bool kMerge = CG == 2 && NPH == 2 && COUT == 128 && KC == 2;   // dw = 0 views as one N = 256 MMA per k16
    const uint32_t cta_rank = CG > 1 ? cluster_ctarank() : 0;
    const bool leader = cta_rank == 0;
    const int unit0 = blockIdx.x / CG, unit_stride = gridDim.x / CG;   // a unit = CG tiles (one per CTA of the pair)
    static_assert(!FUSE || COUT == 64, "fused output conv needs Cout == 64");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_buf = smem;
    uint8_t* b_buf = a_buf + Cfg::kAStages * Cfg::kAStage;
    uint8_t* y_tile = b_buf + Cfg::kBStages * Cfg::kBStage;           // FUSE: [128 x 64] 16-bit, 16 KB
    uint8_t* w4_tile = y_tile + (FUSE ? kATile : 0);                  // FUSE: [32 x 64] 16-bit, 4 KB
    float* s_bias = reinterpret_cast<float*>(w4_tile + (FUSE ? 4096 : 0));
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + COUT);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + Cfg::kAStages;
    uint64_t* b_full = a_empty + Cfg::kAStages;
    uint64_t* b_empty = b_full + Cfg::kBStages;
    uint64_t* acc_full = b_empty + Cfg::kBStages;
    uint64_t* acc_empty = acc_full + 2;
    uint64_t* p_full = acc_empty + 2;
    uint64_t* y_ready = p_full + 1;          // FUSE on CTA pairs, leader: both CTAs' activation tiles are in shared memory
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_ready + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunk = args.Cin / 64;
    const int n_hblk = args.H / args.Hb;
    const int n_units = args.n_tiles / CG * args.pass_split;   // host pads the sample-block count to a multiple of CG
    // small launches: the 8 / NPH passes (output phases) of a tile are spread over pass_split CTAs (pairs) to cut latency
    const int pass_per = (8 / NPH) / args.pass_split;

    for (int i = threadIdx.x; i < COUT; i += kThreads) s_bias[i] = args.bias[i];
    if (FUSE) {
        // CG == 2: this CTA holds tap rows rank*16 .. rank*16+15 of the [32 x 64] tile (N = 32 split over the pair)
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(args.w4tile) + cta_rank * (4096 / CG));
        uint4* dst = reinterpret_cast<uint4*>(w4_tile);
        for (int i = threadIdx.x; i < 4096 / CG / 16; i += kThreads) dst[i] = src[i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (MMA)
    }
    if (threadIdx.x == 0) {
        mbar_init(p_full, 1);
        mbar_init(y_ready, CG);
        for (int i = 0; i < Cfg::kAStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < Cfg::kBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        // CG == 2: one arrive per CTA (elected epilogue thread after the group's named barrier)
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], CG == 1 ? 128 : CG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG > 1) cluster_sync_all();          // peers' barriers are initialised before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= A producer (activation tiles, TMA 5-D box loads) =================
        {
            uint32_t ai = 0;
            uint32_t free_next = mbar_test(&a_empty[0], 1);
            for (int unit = unit0; unit < n_units; unit += unit_stride) {
                // the CTAs of a pair share (t, h block) and take sample blocks CG*k + rank (may be past B: TMA zero-fills, epilogue masks)
                const int tu = unit / args.pass_split, pass0 = (unit % args.pass_split) * pass_per;
                const int hblk = tu % n_hblk, t = (tu / n_hblk) % args.T, bblk = tu / (n_hblk * args.T) * CG + (int)cta_rank;
                const int h0 = hblk * args.Hb, b0 = bblk * args.Bt;
                for (int pass = pass0; pass < pass0 + pass_per; ++pass) {
                    for_each_step<NPH, kSkip, kMerge>(
                        pass, t, args.T, nchunk, KC,
                        [&](int dt, int dh, int dw, int c, bool oob) {
                            if (oob) return;                  // tile entirely in the zero padding: nothing to load
                            const uint32_t s = ai % Cfg::kAStages, ph = (ai / Cfg::kAStages) & 1;
                            if (!free_next) mbar_wait(&a_empty[s], ph ^ 1);
                            ++ai;
                            const uint32_t s2 = ai % Cfg::kAStages, ph2 = (ai / Cfg::kAStages) & 1;
                            free_next = mbar_test(&a_empty[s2], ph2 ^ 1);
                            if (elect_one()) {
                                if (leader) mbar_expect_tx(&a_full[s], CG * Cfg::kAStage);      // both CTAs' tiles
#pragma unroll
                                for (int k = 0; k < KC; ++k) {
                                    if (CG == 1) tma_load_5d(a_buf + s * Cfg::kAStage + k * kATile, &tmap, &a_full[s], (c + k) * 64, dw,
                                                             h0 + dh, t + dt, b0);
                                    else tma_load_5d_cg2(a_buf + s * Cfg::kAStage + k * kATile, &tmap, mapa_rank(smem_u32(&a_full[s]), 0),
                                                         (c + k) * 64, dw, h0 + dh, t + dt, b0);
                                }
                            }
                        },
                        [&](int, int, bool, int, int) {});
                }
            }
        }
    } else if (warp == 2) {
        // ================= B producer (pre-swizzled weight tiles, 1-D bulk copies) =================
        {
            uint32_t bi = 0;
            uint32_t free_next = mbar_test(&b_empty[0], 1);
            for (int unit = unit0; unit < n_units; unit += unit_stride) {
                const int tu = unit / args.pass_split, pass0 = (unit % args.pass_split) * pass_per;
                const int t = (tu / n_hblk) % args.T;
                for (int pass = pass0; pass < pass0 + pass_per; ++pass) {
                    for_each_step<NPH, kSkip, kMerge>(
                        pass, t, args.T, nchunk, KC, [&](int, int, int, int, bool) {},
                        [&](int, int wtile, bool, int mk, int wtile2) {
                            // b_empty[s] completes once the MMA issuers of all CL CTAs released the stage
                            const uint32_t s = bi % Cfg::kBStages, ph = (bi / Cfg::kBStages) & 1;
                            if (!free_next) mbar_wait(&b_empty[s], ph ^ 1);
                            ++bi;
                            const uint32_t s2 = bi % Cfg::kBStages, ph2 = (bi / Cfg::kBStages) & 1;
                            free_next = mbar_test(&b_empty[s2], ph2 ^ 1);
                            if (elect_one()) {
                                if (leader) mbar_expect_tx(&b_full[s], CG * Cfg::kBStage);
                                if (CG == 1) {
                                    const uint8_t* src = reinterpret_cast<const uint8_t*>(args.wpack) + (size_t)wtile * Cfg::kBTile;
                                    bulk_load_1d(b_buf + s * Cfg::kBStage, src, Cfg::kBStage, &b_full[s]);
                                } else {
                                    // weight image as rows of 128 B, COUT rows per tile: this CTA takes rows rank*COUT/2 .. of each tile
                                    const uint32_t bar = mapa_rank(smem_u32(&b_full[s]), 0);
                                    if (kMerge && mk >= 0) {
                                        // merged step: this CTA holds ALL rows of its own phase's tile (rows rank*Cout.. of the N = 2*Cout operand)
                                        const int row0 = (cta_rank == 0 ? wtile : wtile2) * COUT;
#pragma unroll
                                        for (int k = 0; k < 2; ++k)
                                            tma_load_2d_cg2(b_buf + s * Cfg::kBStage + k * (Cfg::kBStage / 2), &tmap_w, bar, 0, row0 + k * (COUT / 2));
                                    } else {
#pragma unroll
                                        for (int k = 0; k < KC; ++k)
                                            tma_load_2d_cg2(b_buf + s * Cfg::kBStage + k * Cfg::kBTile, &tmap_w, bar, 0,
                                                            (wtile + k) * COUT + (int)cta_rank * (COUT / 2));
                                    }
                                }
                            }
                        });
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp walks the schedule; one elected lane issues; leader CTA only) =================
        if (leader) {
            constexpr uint32_t idesc = (1u << 4) | (HalfOps<HT>::kFmt << 7) | (HalfOps<HT>::kFmt << 10) |
                                       ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
            uint32_t ai = 0, bi = 0, acc_it = 0;
            uint32_t a_ready = 0, b_ready = 0;       // results of early probes of the next full barriers
            for (int unit = unit0; unit < n_units; unit += unit_stride) {
                const int tu = unit / args.pass_split, pass0 = (unit % args.pass_split) * pass_per;
                const int t = (tu / n_hblk) % args.T;
                for (int pass = pass0; pass < pass0 + pass_per; ++pass, ++acc_it) {
                    const uint32_t as = acc_it % Cfg::kAccStages, aph = (acc_it / Cfg::kAccStages) & 1;
                    if (CG == 1) mbar_wait(&acc_empty[as], aph ^ 1);
                    else mbar_wait_cluster(&acc_empty[as], aph ^ 1);
                    tc_fence_after();
                    const uint32_t d_base = tmem_base + as * Cfg::kAccCols;
                    uint32_t started = 0;
                    uint32_t cur_a = 0;
                    bool have_a = false;
                    uint32_t prev_a_slot = 0;
                    for_each_step<NPH, kSkip, kMerge>(
                        pass, t, args.T, nchunk, KC,
                        [&](int, int, int, int, bool oob) {
                            if (oob) return;
                            if (have_a && elect_one()) tc_commit_g<CG>(&a_empty[prev_a_slot]);   // MMAs reading the previous A stage issued
                            const uint32_t s = ai % Cfg::kAStages, ph = (ai / Cfg::kAStages) & 1;
                            if (!a_ready) mbar_wait(&a_full[s], ph);
                            ++ai;
                            a_ready = mbar_test(&a_full[ai % Cfg::kAStages], (ai / Cfg::kAStages) & 1);
                            tc_fence_after();
                            cur_a = smem_u32(a_buf + s * Cfg::kAStage);
                            prev_a_slot = s;
                            have_a = true;
                        },
                        [&](int slot, int, bool oob, int mk, int) {
                            const uint32_t s = bi % Cfg::kBStages, ph = (bi / Cfg::kBStages) & 1;
                            if (!b_ready) mbar_wait(&b_full[s], ph);
                            ++bi;
                            b_ready = mbar_test(&b_full[bi % Cfg::kBStages], (bi / Cfg::kBStages) & 1);
                            tc_fence_after();
                            const uint32_t b_addr = smem_u32(b_buf + s * Cfg::kBStage);
                            const uint32_t acc0 = (started >> slot) & 1u;
                            if (elect_one()) {
                                if (kMerge && mk >= 0) {
                                    if (!oob) {
                                        constexpr uint32_t idesc2 = (1u << 4) | (HalfOps<HT>::kFmt << 7) | (HalfOps<HT>::kFmt << 10) |
                                                                    ((uint32_t)((2 * COUT) >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
                                        const uint64_t ad = make_sdesc(cur_a + mk * kATile), bd = make_sdesc(b_addr);
#pragma unroll
                                        for (int k = 0; k < 4; ++k)
                                            tc_mma_g<CG>(d_base, ad + 2 * k, bd + 2 * k, idesc2, (started & 1u) | ((mk | k) ? 1u : 0u));
                                    }
                                } else if (!oob) {
#pragma unroll
                                    for (int c = 0; c < KC; ++c) {
                                        const uint64_t ad = make_sdesc(cur_a + c * kATile), bd = make_sdesc(b_addr + c * Cfg::kBTile);
#pragma unroll
                                        for (int k = 0; k < 4; ++k)   // +32 B per K=16 step inside the 128 B swizzle atom (encoded >>4)
                                            tc_mma_g<CG>(d_base + slot * COUT, ad + 2 * k, bd + 2 * k, idesc, acc0 | ((c | k) ? 1u : 0u));
                                    }
                                }
                                tc_commit_g<CG>(&b_empty[s]);       // CG == 2: releases the stage in both CTAs
                            }
                            if (!oob) started |= (kMerge && mk >= 0) ? 3u : (1u << slot);
                        });
                    if (elect_one()) {
                        if (have_a) tc_commit_g<CG>(&a_empty[prev_a_slot]);
                        tc_commit_g<CG>(&acc_full[as]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ================= epilogue =================
        const int q = warp & 3;                 // TMEM lane quarter this warp may touch (warps 3..6 -> 3,0,1,2)
        const int r = q * 32 + lane;            // accumulator row = position within the tile
        const int bl = r / (args.Hb * args.W), hl = (r / args.W) % args.Hb, w = r % args.W;
        HT* out = reinterpret_cast<HT*>(args.out);
        uint32_t acc_it = 0, p_it = 0;
        for (int unit = unit0; unit < n_units; unit += unit_stride) {
            const int tu = unit / args.pass_split, pass0 = (unit % args.pass_split) * pass_per;
            const int hblk = tu % n_hblk, t = (tu / n_hblk) % args.T, bblk = tu / (n_hblk * args.T) * CG + (int)cta_rank;
            const int b = bblk * args.Bt + bl, h = hblk * args.Hb + hl;
            const bool valid = b < args.B;
            for (int pass = pass0; pass < pass0 + pass_per; ++pass, ++acc_it) {
                const uint32_t as = acc_it % Cfg::kAccStages, aph = (acc_it / Cfg::kAccStages) & 1;
                mbar_wait(&acc_full[as], aph);
                tc_fence_after();
#pragma unroll 1
                for (int s = 0; s < NPH; ++s) {
                    const int p = pass * NPH + s;
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + as * Cfg::kAccCols + s * COUT;
                    float ss = 0.f;
#pragma unroll 1
                    for (int c0 = 0; c0 < COUT; c0 += 32) {
                        uint32_t v[32];
                        tc_ld32(taddr + c0, v);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float x = __uint_as_float(v[j]) + s_bias[c0 + j];
                            ss = fmaf(x, x, ss);
                        }
                    }
                    const float inv = 1.0f / sqrtf(ss * (1.0f / COUT) + 1.0e-8f);
                    const size_t o_pos = (((size_t)b * (2 * args.T) + (2 * t + (p >> 2))) * (2 * args.H) +
                                          (2 * h + ((p >> 1) & 1))) * (2 * args.W) + (2 * w + (p & 1));
#pragma unroll 1
                    for (int c0 = 0; c0 < COUT; c0 += 32) {
                        uint32_t v[32];
                        tc_ld32(taddr + c0, v);
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float x0 = (__uint_as_float(v[2 * j]) + s_bias[c0 + 2 * j]) * inv;
                            float x1 = (__uint_as_float(v[2 * j + 1]) + s_bias[c0 + 2 * j + 1]) * inv;
                            x0 = x0 > 0.f ? x0 : 0.2f * x0;
                            x1 = x1 > 0.f ? x1 : 0.2f * x1;
                            pk[j] = HalfOps<HT>::pack(x0, x1);
                        }
                        if (FUSE) {
                            // row r of the K-major SWIZZLE_128B tile: 16-byte chunk j lives at chunk (j ^ (r & 7))
                            uint8_t* yrow = y_tile + r * 128;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                *reinterpret_cast<uint4*>(yrow + ((((c0 >> 3) + j) ^ (r & 7)) << 4)) =
                                    make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                        } else if (valid) {
                            uint4* dst = reinterpret_cast<uint4*>(out + o_pos * COUT + c0);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                        }
                    }
                    if (FUSE) {
                        // all 128 rows of y written and all reads of this slot's accumulator done
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        tc_fence_before();
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        if (warp == 3) {
                            if (CG > 1) {
                                if (elect_one()) mbar_arrive_cluster(mapa_rank(smem_u32(y_ready), 0));
                                __syncwarp();
                                if (leader) mbar_wait_cluster(y_ready, p_it & 1);
                            }
                            if (leader && elect_one()) {
                                tc_fence_after();
                                constexpr uint32_t idesc_p = (1u << 4) | (HalfOps<HT>::kFmt << 7) | (HalfOps<HT>::kFmt << 10) |
                                                             ((uint32_t)(32 >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
                                const uint64_t yd = make_sdesc(smem_u32(y_tile)), wd = make_sdesc(smem_u32(w4_tile));
                                const uint32_t d_p = tmem_base + as * Cfg::kAccCols + s * COUT;   // reuse the consumed accumulator
#pragma unroll
                                for (int k = 0; k < 4; ++k) tc_mma_g<CG>(d_p, yd + 2 * k, wd + 2 * k, idesc_p, k > 0 ? 1u : 0u);
                                tc_commit_g<CG>(p_full);
                            }
                            __syncwarp();
                        }
                        mbar_wait(p_full, p_it & 1);
                        ++p_it;
                        tc_fence_after();
                        uint32_t pv[32];
                        tc_ld32(taddr, pv);
                        if (valid) {
                            uint4* dst = reinterpret_cast<uint4*>(args.p_out + o_pos * 32);
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                dst[j] = make_uint4(pv[4 * j], pv[4 * j + 1], pv[4 * j + 2], pv[4 * j + 3]);
                        }
                    }
                }
                tc_fence_before();
                if (CG == 1) {
                    mbar_arrive(&acc_empty[as]);
                } else {
                    asm volatile("bar.sync 1, 128;" ::: "memory");          // the four epilogue warps have drained the accumulator
                    if (warp == 3 && lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&acc_empty[as]), 0));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CG > 1) cluster_sync_all();          // no CTA leaves while the pair's MMAs / multicasts may still touch it
    if (warp == 1) {
        tc_fence_after();
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------ last conv + softmax
// Conv3D(1,(3,3,3),'same') + Softmax over the 24 hours (+ optional mm rescale) fused:
// gan_train_cwgangp_pixelnorm.py:345-350, raindisagg_gan_pretrained.py:62-64.
// One CTA = one sample x 256 pixels (HB rows x W).  The 24 input hour-planes are streamed through a
// shared-memory slab with a 1-pixel zero halo; lane l of a warp owns channels (2l, 2l+1), 54 weights
// stay in registers; each input plane t contributes to logits t-1, t, t+1 of the warp's 32 pixels,
// kept as rolling per-lane partials; completed logits are transpose-reduced across the warp so that
// lane i ends up owning pixel i's 24 logits, does the softmax in registers and stores coalesced.
template <typename TI> struct In2;
template <> struct In2<float> {
    static __device__ __forceinline__ float2 ld(const float* p) { return *reinterpret_cast<const float2*>(p); }
};
template <> struct In2<__half> {
    static __device__ __forceinline__ float2 ld(const __half* p) { return __half22float2(*reinterpret_cast<const __half2*>(p)); }
};
template <> struct In2<__nv_bfloat16> {
    static __device__ __forceinline__ float2 ld(const __nv_bfloat16* p) {
        return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
    }
};

// reduce 32 per-lane partials (one per pixel) so that lane i returns the warp-wide sum of v[i]
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            float send = upper ? v[i] : v[i + half];
            float keep = upper ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    return v[0];
}

template <typename TI>
__global__ void __launch_bounds__(256, 1)
conv_out_softmax_kernel(const TI* __restrict__ x, const float* __restrict__ w4, const float* __restrict__ b4,
                        float* __restrict__ out, const float* __restrict__ cond, int B, int lognd, int HB, int spc,
                        int b_off, int ncond, float scale, int out_mm, int* __restrict__ nonfinite) {
    extern __shared__ __align__(16) uint8_t smem_raw2[];
    float* logit_s = reinterpret_cast<float*>(smem_raw2);                       // [24][256]
    TI* slab = reinterpret_cast<TI*>(smem_raw2 + RDG_NHOURS * 256 * 4);         // [(HB+2)][(nd+2)][64]
    const int nd = 1 << lognd;
    const int W2 = nd + 2;
    const int n_hblk = nd / HB;
    const int b = blockIdx.x / n_hblk, h0 = (blockIdx.x % n_hblk) * HB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // pixel i of this warp: index warp*32+i within the HB x nd tile
    float wreg[27][2];
#pragma unroll
    for (int k = 0; k < 27; ++k) { wreg[k][0] = w4[k * 64 + 2 * lane]; wreg[k][1] = w4[k * 64 + 2 * lane + 1]; }

    // zero the slab once (halo stays zero; interior is overwritten per plane)
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(slab);
        const int nwords = (HB + 2) * W2 * 64 * (int)sizeof(TI) / 4;
        for (int i = threadIdx.x; i < nwords; i += 256) z[i] = 0u;
    }
    float accA[32], accB[32], accC[32];   // partial logits for hours t-1, t, t+1 (relative to current plane)
#pragma unroll
    for (int i = 0; i < 32; ++i) { accA[i] = 0.f; accB[i] = 0.f; accC[i] = 0.f; }
    const TI* xb = x + (size_t)b * RDG_NHOURS * nd * nd * 64;
    constexpr int VEC = 16 / (int)sizeof(TI);            // elements per 16-byte vector
    const int vec_per_row = nd * 64 / VEC;               // contiguous (w,c) run of one h row

#pragma unroll 1
    for (int t = 0; t < RDG_NHOURS; ++t) {
        __syncthreads();
        // load rows h0-1 .. h0+HB of plane t (rows outside the domain stay/are zero)
        for (int rr = 0; rr < HB + 2; ++rr) {
            const int h = h0 + rr - 1;
            uint4* drow = reinterpret_cast<uint4*>(slab + ((size_t)rr * W2 + 1) * 64);
            if (h >= 0 && h < nd) {
                const uint4* srow = reinterpret_cast<const uint4*>(xb + ((size_t)t * nd + h) * nd * 64);
                for (int i = threadIdx.x; i < vec_per_row; i += 256) drow[i] = srow[i];
            } else if (t == 0 || n_hblk > 1) {
                for (int i = threadIdx.x; i < vec_per_row; i += 256) drow[i] = make_uint4(0, 0, 0, 0);
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int pix = warp * 32 + i;
            const int ph = pix >> lognd, pw = pix & (nd - 1);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const float2 v = In2<TI>::ld(slab + ((size_t)(ph + kh) * W2 + (pw + kw)) * 64 + 2 * lane);
                    // plane t feeds output hour (t - kt + 1) with tap kt
                    a0 = fmaf(v.x, wreg[(2 * 3 + kh) * 3 + kw][0], fmaf(v.y, wreg[(2 * 3 + kh) * 3 + kw][1], a0));  // hour t-1
                    a1 = fmaf(v.x, wreg[(1 * 3 + kh) * 3 + kw][0], fmaf(v.y, wreg[(1 * 3 + kh) * 3 + kw][1], a1));  // hour t
                    a2 = fmaf(v.x, wreg[(0 * 3 + kh) * 3 + kw][0], fmaf(v.y, wreg[(0 * 3 + kh) * 3 + kw][1], a2));  // hour t+1
                }
            accA[i] += a0; accB[i] += a1; accC[i] += a2;
        }
        // hour t-1 is complete now
        if (t >= 1) logit_s[(t - 1) * 256 + warp * 32 + lane] = transpose_reduce32(accA, lane);
#pragma unroll
        for (int i = 0; i < 32; ++i) { accA[i] = accB[i]; accB[i] = accC[i]; accC[i] = 0.f; }
    }
    logit_s[(RDG_NHOURS - 1) * 256 + warp * 32 + lane] = transpose_reduce32(accA, lane);

    // softmax over hours for pixel `lane` of this warp (each lane re-reads only what it wrote)
    const int pix = warp * 32 + lane;
    const int h = h0 + (pix >> lognd), wq = pix & (nd - 1);
    const float bias = b4[0];
    float logit[RDG_NHOURS];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < RDG_NHOURS; ++t) { logit[t] = logit_s[t * 256 + pix] + bias; mx = fmaxf(mx, logit[t]); }
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < RDG_NHOURS; ++t) { logit[t] = expf(logit[t] - mx); s += logit[t]; }
    float mul = 1.f;
    if (out_mm) mul = cond[(((size_t)((b_off + b) / spc) * nd + h) * nd + wq) * ncond] * scale;
    bool bad = false;
    float* ob = out + (size_t)b * RDG_NHOURS * nd * nd + (size_t)h * nd + wq;
#pragma unroll
    for (int t = 0; t < RDG_NHOURS; ++t) {
        const float f = logit[t] / s;
        bad |= !isfinite(f);
        ob[(size_t)t * nd * nd] = f * mul;
    }
    if (bad && nonfinite) atomicOr(nonfinite, 1);
}

// logit[q] = b + sum_{27 taps} P[q + offset(tap)][tap]; softmax over the 24 hours; optional mm rescale.
// One CTA = one sample x 256 pixels (HB rows x nd); one thread per pixel.  The P planes (hour by hour)
// are staged in shared memory with a zero halo and a 33-word position stride (bank-conflict free);
// input hour t feeds logits t+1, t, t-1 through taps kt = 0, 1, 2.
__global__ void __launch_bounds__(256)
gather_softmax_kernel(const float* __restrict__ P, const float* __restrict__ b4, float* __restrict__ out,
                      const float* __restrict__ cond, int lognd, int HB, int spc, int b_off, int ncond, float scale,
                      int out_mm, int* __restrict__ nonfinite) {
    extern __shared__ __align__(16) float ps[];          // [(HB+2)][(nd+2)][33]
    const int nd = 1 << lognd, W2 = nd + 2;
    const int n_hblk = nd / HB;
    const int b = blockIdx.x / n_hblk, h0 = (blockIdx.x % n_hblk) * HB;
    const int tid = threadIdx.x;
    const int ph = tid >> lognd, pw = tid & (nd - 1);
    for (int i = tid; i < (HB + 2) * W2 * 33; i += 256) ps[i] = 0.f;
    float acc[RDG_NHOURS + 2];
#pragma unroll
    for (int t = 0; t < RDG_NHOURS + 2; ++t) acc[t] = 0.f;
    const float* pb = P + (size_t)b * RDG_NHOURS * nd * nd * 32;
    const int f4_per_row = nd * 8;                       // float4 per h-row (nd positions x 32 floats)
#pragma unroll
    for (int t = 0; t < RDG_NHOURS; ++t) {
        // the HB+2 rows of this hour are one contiguous run of (HB+2)*nd*8 float4 (minus rows outside the
        // domain, which stay zero: zeroed once, never written).  All loads of a thread are issued before its
        // first shared-memory store so their latencies overlap.
        const int r_lo = h0 - 1 < 0 ? 1 : 0, r_hi = h0 + HB + 1 > nd ? HB + 1 : HB + 2;      // valid slab rows [r_lo, r_hi)
        const int n_f4 = (r_hi - r_lo) * f4_per_row;
        const float4* src = reinterpret_cast<const float4*>(pb + ((size_t)t * nd + (h0 - 1 + r_lo)) * nd * 32);
        __syncthreads();
        for (int base = 0; base < n_f4; base += 256 * 8) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = base + k * 256 + tid;
                if (i < n_f4) v[k] = src[i];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = base + k * 256 + tid;
                if (i < n_f4) {
                    const int rr = r_lo + (i >> (lognd + 3)), j = i & (f4_per_row - 1);
                    float* d = ps + ((size_t)rr * W2 + 1 + (j >> 3)) * 33 + (j & 7) * 4;
                    d[0] = v[k].x; d[1] = v[k].y; d[2] = v[k].z; d[3] = v[k].w;
                }
            }
        }
        __syncthreads();
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                // output pixel (ph,pw) with tap (kh,kw) reads input pixel (ph+kh-1, pw+kw-1); slab index has +1 halo
                const float* q = ps + ((size_t)(ph + kh) * W2 + (pw + kw)) * 33;
                s0 += q[(0 * 3 + kh) * 3 + kw];          // kt = 0: this plane is hour (t_out - 1) -> t_out = t + 1
                s1 += q[(1 * 3 + kh) * 3 + kw];          // kt = 1: t_out = t
                s2 += q[(2 * 3 + kh) * 3 + kw];          // kt = 2: t_out = t - 1
            }
        acc[t + 2] += s0; acc[t + 1] += s1; acc[t] += s2;   // acc index = t_out + 1
    }
    const int h = h0 + ph;
    const float bias = b4[0];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < RDG_NHOURS; ++t) { acc[t + 1] += bias; mx = fmaxf(mx, acc[t + 1]); }
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < RDG_NHOURS; ++t) { acc[t + 1] = expf(acc[t + 1] - mx); s += acc[t + 1]; }
    float mul = 1.f;
    if (out_mm) mul = cond[(((size_t)((b_off + b) / spc) * nd + h) * nd + pw) * ncond] * scale;
    bool bad = false;
    float* ob = out + (size_t)b * RDG_NHOURS * nd * nd + (size_t)h * nd + pw;
#pragma unroll
    for (int t = 0; t < RDG_NHOURS; ++t) {
        const float f = acc[t + 1] / s;
        bad |= !isfinite(f);
        ob[(size_t)t * nd * nd] = f * mul;
    }
    if (bad && nonfinite) atomicOr(nonfinite, 1);
}

template <typename HT>
__global__ void f32_to_half_kernel(const float* __restrict__ src, HT* __restrict__ dst, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = HalfOps<HT>::from_float(src[i]);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

template <typename HT, int COUT, int NPH, int KC, bool FUSE, int CG>
int launch_upconv_cg(const void* x, const void* wpack, const float* bias, void* y, const void* w4tile, float* p_out, int B,
                  int T, int H, int W, int Cin, int sm_count, cudaStream_t st) {
    using Cfg = TcCfg<COUT, NPH, KC, CG>;
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { rdg_set_error("cuTensorMapEncodeTiled entry point not available"); return RDG_TC_E_DRIVER; }
    if (Cin % (64 * KC) || W > 128 || (128 % W) != 0) { rdg_set_error("tc upconv: unsupported shape"); return RDG_TC_E_SHAPE; }
    TcConvArgs a;
    a.B = B; a.T = T; a.H = H; a.W = W; a.Cin = Cin;
    int rows_per_sample_plane = H * W;
    if (rows_per_sample_plane >= 128) { a.Hb = 128 / W; a.Bt = 1; }
    else { a.Hb = H; a.Bt = 128 / rows_per_sample_plane; }
    if (a.Hb < 1 || H % a.Hb) { rdg_set_error("tc upconv: H not divisible by tile rows"); return RDG_TC_E_SHAPE; }
    a.n_tiles = ceil_div(ceil_div(B, a.Bt), CG) * CG * T * (H / a.Hb);    // sample blocks padded to a multiple of CG
    a.wpack = wpack; a.bias = bias; a.out = y; a.w4tile = w4tile; a.p_out = p_out;

    CUtensorMap tmap;
    cuuint64_t gdim[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t gstr[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2,
                          (cuuint64_t)T * H * W * Cin * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)W, (cuuint32_t)a.Hb, 1, (cuuint32_t)a.Bt};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUtensorMapDataType dt = sizeof(HT) == 2 && HalfOps<HT>::kFmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                                       : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = enc(&tmap, dt, 5, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { rdg_set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return RDG_TC_E_DRIVER; }

    // folded weight image as rows of 128 B (COUT rows per (phase, tap, chunk) tile); box = one CTA's half tile (CG == 2)
    CUtensorMap tmap_w;
    {
        cuuint64_t wdim[2] = {64, (cuuint64_t)64 * (Cin / 64) * COUT};
        cuuint64_t wstr[1] = {128};
        cuuint32_t wbox[2] = {64, (cuuint32_t)(COUT / 2)};
        cuuint32_t wes[2] = {1, 1};
        r = enc(&tmap_w, dt, 2, const_cast<void*>(wpack), wdim, wstr, wbox, wes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { rdg_set_error("cuTensorMapEncodeTiled (weights) failed: %d", (int)r); return RDG_TC_E_DRIVER; }
    }
    auto kern = tc_upconv_pixelnorm_kernel<HT, COUT, NPH, KC, FUSE, CG>;
    static bool attr_set = false;
    if (!attr_set) {
        RDG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
        attr_set = true;
    }
    const int max_clusters = sm_count / CG;
    a.pass_split = 1;
    while (a.pass_split * 2 <= 8 / NPH && a.n_tiles / CG * a.pass_split * 2 <= max_clusters) a.pass_split *= 2;
    const int units = a.n_tiles / CG * a.pass_split;
    const int grid = (units < max_clusters ? units : max_clusters) * CG;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = Cfg::kSmem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    RDG_CUDA(cudaLaunchKernelEx(&cfg, kern, tmap, tmap_w, a));
    return 0;
}

template <typename HT, int COUT, int NPH, int KC, bool FUSE>
int launch_upconv(const void* x, const void* wpack, const float* bias, void* y, const void* w4tile, float* p_out, int B,
                  int T, int H, int W, int Cin, int sm_count, cudaStream_t st) {
    // CTA pairs (cta_group::2); RDG_CG=1 forces single-CTA MMAs.
    // (An earlier weight-multicast cluster variant of the single-CTA kernel measured no gain and was removed.)
    static const int cg = getenv("RDG_CG") ? atoi(getenv("RDG_CG")) : 2;
    if (cg == 2)
        return launch_upconv_cg<HT, COUT, NPH, KC, FUSE, 2>(x, wpack, bias, y, w4tile, p_out, B, T, H, W, Cin, sm_count, st);
    return launch_upconv_cg<HT, COUT, NPH, KC, FUSE, 1>(x, wpack, bias, y, w4tile, p_out, B, T, H, W, Cin, sm_count, st);
}

template <typename HT>
int tc_upconv_dispatch(const void* x, const void* wpack, const float* bias, void* y, const void* w4tile, float* p_out,
                       int B, int T, int H, int W, int Cin, int Cout, int sm_count, cudaStream_t st) {
    // phases per pass (NPH): more phases share each activation stage (fewer L2 re-reads) but NPH*Cout > 256
    // TMEM columns forfeits the double-buffered accumulator (no MMA/epilogue overlap).  Tunable for experiments.
    static const int nph64 = getenv("RDG_NPH64") ? atoi(getenv("RDG_NPH64")) : 4;
    static const int nph128 = getenv("RDG_NPH128") ? atoi(getenv("RDG_NPH128")) : 2;
    static const int nph256 = getenv("RDG_NPH256") ? atoi(getenv("RDG_NPH256")) : 1;
    if (Cout == 256) {
        if (nph256 == 2) return launch_upconv<HT, 256, 2, 1, false>(x, wpack, bias, y, nullptr, nullptr, B, T, H, W, Cin, sm_count, st);
        return launch_upconv<HT, 256, 1, 1, false>(x, wpack, bias, y, nullptr, nullptr, B, T, H, W, Cin, sm_count, st);
    }
    if (Cout == 128) {
        if (nph128 == 4) return launch_upconv<HT, 128, 4, 2, false>(x, wpack, bias, y, nullptr, nullptr, B, T, H, W, Cin, sm_count, st);
        return launch_upconv<HT, 128, 2, 2, false>(x, wpack, bias, y, nullptr, nullptr, B, T, H, W, Cin, sm_count, st);
    }
    if (Cout == 64 && p_out) {
        if (nph64 == 8) return launch_upconv<HT, 64, 8, 2, true>(x, wpack, bias, y, w4tile, p_out, B, T, H, W, Cin, sm_count, st);
        if (nph64 == 2) return launch_upconv<HT, 64, 2, 2, true>(x, wpack, bias, y, w4tile, p_out, B, T, H, W, Cin, sm_count, st);
        return launch_upconv<HT, 64, 4, 2, true>(x, wpack, bias, y, w4tile, p_out, B, T, H, W, Cin, sm_count, st);
    }
    if (Cout == 64) return launch_upconv<HT, 64, 4, 2, false>(x, wpack, bias, y, nullptr, nullptr, B, T, H, W, Cin, sm_count, st);
    rdg_set_error("tc upconv: unsupported Cout %d", Cout);
    return RDG_TC_E_SHAPE;
}

template <typename TI>
int launch_conv_out(const void* x, const float* w4, const float* b4, float* out, const float* cond, int B, int nd,
                    int spc, int b_off, int ncond, float scale, int out_mm, int* nonfinite, cudaStream_t st) {
    int lognd = 0;
    while ((1 << lognd) < nd) ++lognd;
    if ((1 << lognd) != nd || nd > 256) { rdg_set_error("conv_out: nd must be a power of two <= 256"); return RDG_TC_E_SHAPE; }
    int HB = 256 / nd;
    if (HB > nd) { rdg_set_error("conv_out: nd too small"); return RDG_TC_E_SHAPE; }
    size_t smem = (size_t)(HB + 2) * (nd + 2) * 64 * sizeof(TI) + RDG_NHOURS * 256 * 4;
    auto kern = conv_out_softmax_kernel<TI>;
    static bool attr_set = false;
    if (!attr_set) {
        RDG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    if (smem > 200 * 1024) { rdg_set_error("conv_out: slab too large"); return RDG_TC_E_SHAPE; }
    int grid = B * (nd / HB);
    kern<<<grid, 256, smem, st>>>(reinterpret_cast<const TI*>(x), w4, b4, out, cond, B, lognd, HB, spc, b_off, ncond, scale,
                                  out_mm, nonfinite);
    RDG_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int tc_upconv_pixelnorm(int half_kind, const void* x, const void* wpack, const float* bias, void* y, const void* w4tile,
                        float* p_out, int B, int T, int H, int W, int Cin, int Cout, int sm_count, cudaStream_t st) {
    if (B <= 0) return 0;
    if (half_kind == RDG_HALF_BF16)
        return tc_upconv_dispatch<__nv_bfloat16>(x, wpack, bias, y, w4tile, p_out, B, T, H, W, Cin, Cout, sm_count, st);
    return tc_upconv_dispatch<__half>(x, wpack, bias, y, w4tile, p_out, B, T, H, W, Cin, Cout, sm_count, st);
}

// ------------------------------------------------------------------ weight packing (device side)
namespace {
// One thread per packed element.  Folds the nearest-x2 upsample into the 3^3 kernel:
// per axis, phase 0: tap0 <- {k0}, tap1 <- {k1,k2}; phase 1: tap0 <- {k0,k1}, tap1 <- {k2}; sums in f32,
// rounded once to 16 bit, stored as [phase][tap][chunk][Cout rows x 64 k] with the 16-byte chunks of
// row n XOR-swizzled by (n & 7) -- the SWIZZLE_128B K-major image tcgen05.mma reads.
// A block packs one (phase, tap, 64-channel chunk, 64 output channels) sub-tile through shared memory: the source is read along
// Cout (its fast index), the image written along the 64 k positions of a row (its fast index).  The training loop re-packs after
// every generator update; one thread per packed element read the source with a stride of Cout floats (56 us for the 256 -> 256
// layer, now a few).  Same sums in the same order as before (kt, kh, kw ascending), rounded once.
template <typename HT>
__global__ void __launch_bounds__(256) pack_folded_kernel(const float* __restrict__ k, HT* __restrict__ dst, int Cin, int Cout) {
    __shared__ float tile[64][65];
    const int nchunk = Cin / 64, nblk_n = Cout / 64;
    int b = blockIdx.x;
    const int nb = b % nblk_n; b /= nblk_n;
    const int chunk = b % nchunk; b /= nchunk;
    const int a = b % 8, p = b / 8;
    const int pt = p >> 2, ph = (p >> 1) & 1, pw = p & 1, at = a >> 2, ah = (a >> 1) & 1, aw = a & 1;
    auto lo = [](int phs, int tp) { return phs == 0 ? (tp == 0 ? 0 : 1) : (tp == 0 ? 0 : 2); };
    auto hi = [](int phs, int tp) { return phs == 0 ? (tp == 0 ? 0 : 2) : (tp == 0 ? 1 : 2); };
    const int nl = threadIdx.x & 63, r0 = threadIdx.x >> 6;
    for (int cl = r0; cl < 64; cl += 4) {
        const int ci = chunk * 64 + cl;
        float sum = 0.f;
        for (int kt = lo(pt, at); kt <= hi(pt, at); ++kt)
            for (int kh = lo(ph, ah); kh <= hi(ph, ah); ++kh)
                for (int kw = lo(pw, aw); kw <= hi(pw, aw); ++kw)
                    sum += k[((long long)((kt * 3 + kh) * 3 + kw) * Cin + ci) * Cout + nb * 64 + nl];
        tile[cl][nl] = sum;
    }
    __syncthreads();
    // image: [phase][tap][chunk][Cout rows x 64 k], 16-byte chunks of row n XOR-swizzled by (n & 7)
    HT* out = dst + ((long long)((p * 8 + a) * nchunk + chunk) * Cout + nb * 64) * 64;
    const int pos = threadIdx.x & 63;
    for (int n_l = r0; n_l < 64; n_l += 4) {
        const int n = nb * 64 + n_l;
        const int j = (pos >> 3) ^ (n & 7), e = pos & 7;
        out[n_l * 64 + pos] = HalfOps<HT>::from_float(tile[j * 8 + e][n_l]);
    }
}
// output conv (3,3,3,64,1) as a [32 taps x 64 ch] swizzled B tile (taps 27..31 zero)
template <typename HT>
__global__ void pack_w4_kernel(const float* __restrict__ k4, HT* __restrict__ dst) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 32 * 64) return;
    const int n = idx / 64, pos = idx % 64;
    const int j = (pos >> 3) ^ (n & 7), e = pos & 7;
    dst[idx] = HalfOps<HT>::from_float(n < 27 ? k4[n * 64 + j * 8 + e] : 0.f);
}
// Fixed-point scale of the on-chip output-conv sum (gen_tc_planes.cu): |sum_k P_k| <= sum_k |y|_2 |w_k|_2 and
// PixelNorm bounds |y|_2 <= sqrt(64) = 8, so 2 * 8 * sum_k |w_k|_2 (2x margin for the 16-bit roundings) bounds every
// partial sum; scale = largest power of two keeping that below 2^31.
__global__ void __launch_bounds__(27 * 32) w4_scale_kernel(const float* __restrict__ k4, float* __restrict__ dst) {
    __shared__ float nrm[27];
    const int tap = threadIdx.x >> 5, lane = threadIdx.x & 31;       // one warp per tap
    const float a = k4[tap * 64 + lane], b = k4[tap * 64 + 32 + lane];
    float q = a * a + b * b;
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) nrm[tap] = sqrtf(q);
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int t = 0; t < 27; ++t) s += nrm[t];
        const float bound = 16.f * s;
        int e = 20;
        if (bound > 0.f && bound < INFINITY) {
            e = 30 - (int)ceilf(log2f(bound));
            e = e > 30 ? 30 : (e < -60 ? -60 : e);
        }
        dst[0] = exp2f((float)e);
        dst[1] = exp2f((float)-e);
    }
}

// One thread per (sample, pixel): the 24 fixed-point logits of the pixel -> softmax over the hours, in place.
__global__ void __launch_bounds__(256)
softmax_fixed_kernel(float* __restrict__ out, const float* __restrict__ scales, const float* __restrict__ b4,
                     const float* __restrict__ cond, int spc, int b_off, int ncond, float norm_scale, int out_mm,
                     int* __restrict__ nonfinite) {
    const int b = blockIdx.x, px = threadIdx.x;
    const float inv = scales[1], bias = b4[0];
    float* ob = out + (size_t)b * RDG_NHOURS * 256 + px;
    float v[RDG_NHOURS];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < RDG_NHOURS; ++t) {
        v[t] = (float)__float_as_int(ob[t * 256]) * inv + bias;
        mx = fmaxf(mx, v[t]);
    }
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < RDG_NHOURS; ++t) { v[t] = expf(v[t] - mx); s += v[t]; }
    float mul = 1.f;
    if (out_mm) mul = cond[((size_t)((b_off + b) / spc) * 256 + px) * ncond] * norm_scale;
    bool bad = false;
#pragma unroll
    for (int t = 0; t < RDG_NHOURS; ++t) {
        const float f = v[t] / s;
        bad |= !isfinite(f);
        ob[t * 256] = f * mul;
    }
    if (bad && nonfinite) atomicOr(nonfinite, 1);
}
}  // namespace

int softmax_fixed_inplace(float* out, const void* w4tile, const float* b4, const float* cond, int B, int spc, int b_off,
                          int ncond, float norm_scale, int out_mm, int* nonfinite, cudaStream_t st) {
    if (B <= 0) return 0;
    const float* scales = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(w4tile) + 32 * 64 * 2);
    softmax_fixed_kernel<<<B, 256, 0, st>>>(out, scales, b4, cond, spc, b_off, ncond, norm_scale, out_mm, nonfinite);
    RDG_LAUNCH_CHECK();
    return 0;
}

int pack_folded_weights(int half_kind, const float* k, void* dst, int Cin, int Cout, cudaStream_t st) {
    if ((Cin & 63) || (Cout & 63)) { rdg_set_error("pack_folded_weights: Cin and Cout must be multiples of 64"); return RDG_TC_E_SHAPE; }
    const int blocks = 64 * (Cin / 64) * (Cout / 64);
    if (half_kind == RDG_HALF_BF16)
        pack_folded_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(k, (__nv_bfloat16*)dst, Cin, Cout);
    else
        pack_folded_kernel<__half><<<blocks, 256, 0, st>>>(k, (__half*)dst, Cin, Cout);
    RDG_LAUNCH_CHECK();
    return 0;
}
int pack_w4_tile(int half_kind, const float* k4, void* dst, cudaStream_t st) {
    if (half_kind == RDG_HALF_BF16) pack_w4_kernel<__nv_bfloat16><<<8, 256, 0, st>>>(k4, (__nv_bfloat16*)dst);
    else pack_w4_kernel<__half><<<8, 256, 0, st>>>(k4, (__half*)dst);
    RDG_LAUNCH_CHECK();
    w4_scale_kernel<<<1, 27 * 32, 0, st>>>(k4, reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(dst) + 32 * 64 * 2));
    RDG_LAUNCH_CHECK();
    return 0;
}

int gather_softmax(const float* p, const float* b4, float* out, const float* cond, int B, int nd, int spc, int b_off,
                   int ncond, float scale, int out_mm, int* nonfinite, cudaStream_t st) {
    if (B <= 0) return 0;
    int lognd = 0;
    while ((1 << lognd) < nd) ++lognd;
    if ((1 << lognd) != nd || nd > 256 || nd < 16) { rdg_set_error("gather_softmax: nd must be a power of two in [16,256]"); return RDG_TC_E_SHAPE; }
    const int HB = 256 / nd;
    const size_t smem = (size_t)(HB + 2) * (nd + 2) * 33 * 4;
    static bool attr_set = false;
    if (!attr_set) {
        RDG_CUDA(cudaFuncSetAttribute(gather_softmax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    gather_softmax_kernel<<<B * (nd / HB), 256, smem, st>>>(p, b4, out, cond, lognd, HB, spc, b_off, ncond, scale, out_mm, nonfinite);
    RDG_LAUNCH_CHECK();
    return 0;
}

int conv_out_softmax(int in_kind, const void* x, const float* w4, const float* b4, float* out, const float* cond,
                     int B, int nd, int spc, int b_off, int ncond, float scale, int out_mm, int* nonfinite, cudaStream_t st) {
    if (B <= 0) return 0;
    if (in_kind == RDG_HALF_BF16)
        return launch_conv_out<__nv_bfloat16>(x, w4, b4, out, cond, B, nd, spc, b_off, ncond, scale, out_mm, nonfinite, st);
    if (in_kind == RDG_HALF_FP16)
        return launch_conv_out<__half>(x, w4, b4, out, cond, B, nd, spc, b_off, ncond, scale, out_mm, nonfinite, st);
    return launch_conv_out<float>(x, w4, b4, out, cond, B, nd, spc, b_off, ncond, scale, out_mm, nonfinite, st);
}

namespace {
template <typename HT>
__global__ void half_to_f32_kernel(const HT* __restrict__ src, float* __restrict__ dst, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (float)src[i];
}
}  // namespace

int half_to_f32(int half_kind, const void* src, float* dst, long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    if (half_kind == RDG_HALF_BF16)
        half_to_f32_kernel<__nv_bfloat16><<<ceil_div(n, 256), 256, 0, st>>>((const __nv_bfloat16*)src, dst, n);
    else
        half_to_f32_kernel<__half><<<ceil_div(n, 256), 256, 0, st>>>((const __half*)src, dst, n);
    RDG_LAUNCH_CHECK();
    return 0;
}

int f32_to_half(int half_kind, const float* src, void* dst, long long n, cudaStream_t st) {
    if (n <= 0) return 0;
    if (half_kind == RDG_HALF_BF16)
        f32_to_half_kernel<__nv_bfloat16><<<ceil_div(n, 256), 256, 0, st>>>(src, (__nv_bfloat16*)dst, n);
    else
        f32_to_half_kernel<__half><<<ceil_div(n, 256), 256, 0, st>>>(src, (__half*)dst, n);
    RDG_LAUNCH_CHECK();
    return 0;
}
