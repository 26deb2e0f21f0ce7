// Resident-plane tensor-core kernel for the dominant generator layer at ndomain = 16:
// UpSampling3D(2) + Conv3D(128 -> 64, 3^3, 'same') + bias + PixelNormalization + LeakyReLU(0.2)
// (gan_train_cwgangp_pixelnorm.py:340-343) with the output Conv3D(64 -> 1) (:345) fused as tap products.
//
// Why a second kernel: the tile-per-offset kernel in gen_tc.cu re-fetches every shifted activation tile
// and every weight tile from L2 (2.2 MB per 128-position tile) and is bound by the L2 -> SM stream
// (profiles/r1c_ncu_conv3_summary.md).  Here
//   * a CTA owns a PAIR of samples and sweeps their 12 low-res hour planes; three planes stay resident in
//     shared memory as zero-haloed images written by ONE TMA box load each ([h' 10][sample 2][w' 9] rows of
//     64 channels, SWIZZLE_128B).  All 27 (dt,dh,dw) operand views of a plane are plain descriptor offsets
//     (start row (1+dh)*18 + (1+dw), 8-row group stride 9 rows): tcgen05 and TMA both swizzle on absolute
//     shared-memory address bits, so views need not be 1024-byte aligned (tools/umma_shift_probe.cu).
//     The zero halos are shared between neighbouring lines / buffers, so a plane costs 40.5 KB, not 64 KB.
//   * two M tiles (hour planes t0, t0+1 of the pair) share every weight stage, and the two w-phases that
//     need the same view (dw = 0) are one N = 128 MMA; the operand stream per 128 positions drops from
//     2.2 MB to 0.55 MB and shared-memory operand reads per MMA from 6 KB/32 clk to <= 4 KB + 2 KB*N/64.
//
// Pass structure: super-tile = (pair, t0 even); pass = (pt, ph) with accumulators [M tile m][pw][64] = 256
// TMEM columns, double buffered; per pass 16 weight stages of 16 KB: for (at, ah, chunk): stage X = the two
// dw = 0 tiles (pw0,aw1 | pw1,aw0) -> one N=128 MMA per k16 and M tile; stage Y = (pw0,aw0) for dw = -1 and
// (pw1,aw1) for dw = +1 -> two N=64 MMAs.
//
// Warp roles (384 threads): 0 = plane producer (TMA), 1 = TMEM alloc + MMA issuer of M tile 0, 2 = weight producer,
// 3 = MMA issuer of M tile 1, 4..7 / 8..11 = epilogue warpgroup of M tile 0 / 1 (bias + PixelNorm + LeakyReLU, fused
// output-conv tap products, stores); each M tile has its own accumulator full/empty barriers, y tile and P barrier.
// Two issuers because ONE thread sustains at most one tcgen05.mma per ~77 clk (tools/umma_rate_probe.cu): a single
// issuer caps N=64 / N=128 MMAs at 42 % / 84 % of the pipe, two reach the SM limits (48 clk smem-bound / 64 clk).
#include "rdg_common.cuh"
#include "gen_tc.h"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <cstdlib>

using namespace rdg_tc;

namespace {

constexpr int kThreads = 384;
constexpr int kRow = 128;                       // bytes per row: 64 channels x 2 B
constexpr int kWp = 9;                          // rows per (h', sample) line: w' = 0..8 <-> w = -1..7; w = 8 is the next line's w' = 0
constexpr int kLine = 2 * kWp;                  // rows per h' line (the two samples interleaved)
constexpr int kHp = 9;                          // h' lines per buffer; line 9 (h = 8) is the next buffer's line 0
constexpr int kBufBytes = kHp * kLine * kRow;   // 20736: one plane, one 64-channel chunk
constexpr int kBoxBytes = 10 * kLine * kRow;    // 23040: what one TMA box writes (10 lines)
constexpr int kSlots = 3;                       // resident planes
constexpr int kChunks = 2;                      // Cin = 128
constexpr int kARegion = ((kSlots * kChunks * kBufBytes + (kLine + 1) * kRow) + 1023) / 1024 * 1024;
constexpr int kBStageFull = 16384;              // two [64 x 64] weight tiles (a CTA of a pair holds half of each)
constexpr int kBBytes = 5 * kBStageFull;        // weight ring: 5 stages (CG = 1) / 10 stages (CG = 2)
constexpr int kStagesPerPass = 16;
constexpr int kW4Tile = 4096;
// Fused-logit mode: rolling ring of 8 output hours, fixed-point int32 logits of the sample pair.
// word = slot * kRingPlane + h * kRingPitch + w * 2 + sample; pitch 33 makes the scatter of a warp
// (lanes = 8 w x 2 samples x 2 h) hit 32 distinct banks.
constexpr int kRingPitch = 33;
constexpr int kRingPlane = 16 * kRingPitch;
constexpr int kRingSlots = 8;
constexpr int kRingBytes = kRingSlots * kRingPlane * 4;
constexpr int kSmem = 1024 + kBBytes + kW4Tile + kARegion + 512 + kRingBytes;
static_assert(kSmem <= 227 * 1024, "shared memory overflow");
constexpr int kAccCols = 256;
constexpr uint32_t kSboA = kWp * kRow;          // 8-row group stride of an activation view

__device__ __forceinline__ uint64_t make_sdesc_sbo(uint32_t saddr, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <typename HT, int CG>
__global__ void __launch_bounds__(kThreads, 1)
tc_upconv64_planes_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_w, TcConvArgs args) {
    constexpr int kBStage = kBStageFull / CG;       // bytes of a weight stage in THIS CTA's shared memory
    constexpr int kBStages = kBBytes / kBStage;
    const uint32_t rank = CG > 1 ? cluster_ctarank() : 0;
    const bool leader = rank == 0;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* b_buf = smem;
    uint8_t* w4_tile = b_buf + kBBytes;
    uint8_t* a_reg = w4_tile + kW4Tile;
    __shared__ float s_bias[64];                 // static: read with LDS, not generic loads
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_reg + kARegion);
    uint64_t* a_full = bars;
    uint64_t* a_empty = a_full + kSlots;
    uint64_t* b_full = a_empty + kSlots;
    uint64_t* b_empty = b_full + kBStages;
    uint64_t* acc_full = b_empty + kBStages;     // [stage 2][M tile 2]
    uint64_t* acc_empty = acc_full + 4;          // [stage 2][M tile 2]
    uint64_t* p_full = acc_empty + 4;            // [M tile 2][w-phase 2]: one commit and one wait per pass each
    uint64_t* y_ready = p_full + 4;              // CG == 2, leader: [M tile 2][w-phase 2], one arrive per CTA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_ready + 4);
    int* ring = reinterpret_cast<int*>(a_reg + kARegion + 512);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = args.T, n_super = T / 2;
    const int n_units = (args.B + 1) / 2;
    const bool logit_mode = args.logit_out != nullptr;          // output conv accumulated on chip (fixed point)
    const bool fuse = args.p_out != nullptr || logit_mode;

    for (int i = threadIdx.x; i < 64; i += kThreads) s_bias[i] = args.bias[i];
    {
        // halo rows shared between buffers (and the row past the last line) must read as zero from the start
        uint4* z = reinterpret_cast<uint4*>(a_reg);
        for (int i = threadIdx.x; i < kARegion / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < kRingSlots * kRingPlane; i += kThreads) ring[i] = 0;
        if (fuse) {
            // CG == 2: this CTA holds tap rows rank*16 .. rank*16+15 of the [32 x 64] tile (N = 32 split over the pair)
            const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(args.w4tile) + rank * (kW4Tile / CG));
            uint4* dst = reinterpret_cast<uint4*>(w4_tile);
            for (int i = threadIdx.x; i < kW4Tile / CG / 16; i += kThreads) dst[i] = src[i];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (TMA, MMA)
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&p_full[i], 1);
        for (int i = 0; i < kSlots; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 2); }     // both issuers release
        for (int i = 0; i < kBStages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 2); }
        // accumulator release: one arrive per CTA (elected thread after the epilogue group's named barrier)
        for (int i = 0; i < 4; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], CG); mbar_init(&y_ready[i], CG); }
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
    if (CG > 1) cluster_sync_all();              // the peer's barriers are initialised before any remote arrive / multicast
    tc_fence_after();
    const int grp0 = blockIdx.x / CG, grp_stride = gridDim.x / CG;   // a group = CG consecutive units, one per CTA of the pair
    const int n_groups = (n_units + CG - 1) / CG;
    // Work item = (group, range of super-tiles).  Large launches: one item per group (all n_super super-tiles).  Small launches
    // (args.sup_split): one item per (group, super-tile), so a few sample pairs still spread over many SMs; an item then loads
    // planes 2sup-1 .. 2sup+2, and in the on-chip-logit mode flushes ALL the hours it touched with global integer atomics
    // (the neighbouring items add the rest; the host zeroes the buffer first).
    const int sup_per = args.sup_split ? 1 : n_super;
    const int items_per_grp = n_super / sup_per, n_items = n_groups * items_per_grp;
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= plane producer: one haloed box per plane and 64-channel chunk =================
        uint32_t li = 0;
        for (int item = grp0; item < n_items; item += grp_stride) {
            const int grp = item / items_per_grp, sup_lo = (item % items_per_grp) * sup_per, sup_hi = sup_lo + sup_per;
            const int p_lo = 2 * sup_lo - 1 < 0 ? 0 : 2 * sup_lo - 1, p_hi = 2 * sup_hi > T - 1 ? T - 1 : 2 * sup_hi;
            const int b0 = (grp * CG + (int)rank) * 2;            // may be >= B (odd tail): TMA zero-fills, epilogue masks
            for (int p = p_lo; p <= p_hi; ++p, ++li) {
                const uint32_t s = li % kSlots, ph = (li / kSlots) & 1;
                mbar_wait(&a_empty[s], ph ^ 1);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(&a_full[s], CG * kChunks * kBoxBytes);      // bytes of both CTAs' planes
#pragma unroll
                    for (int c = 0; c < kChunks; ++c) {  // tensor dims (C, W, B, H, T): w and h start at -1 (zero halo)
                        if (CG == 1) tma_load_5d(a_reg + (s * kChunks + c) * kBufBytes, &tmap, &a_full[s], c * 64, -1, b0, -1, p);
                        else tma_load_5d_cg2(a_reg + (s * kChunks + c) * kBufBytes, &tmap, mapa_rank(smem_u32(&a_full[s]), 0), c * 64, -1, b0, -1, p);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 2) {
        // ================= weight producer: the 1 MB stage sequence of a super-tile, streamed in order =================
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(args.wpack);
        uint32_t bi = 0;
        for (int item = grp0; item < n_items; item += grp_stride)
            for (int sup = 0; sup < sup_per; ++sup)
                for (int st = 0; st < 4 * kStagesPerPass; ++st, ++bi) {
                    const uint32_t s = bi % kBStages, ph = (bi / kBStages) & 1;
                    mbar_wait(&b_empty[s], ph ^ 1);
                    if (elect_one()) {
                        if (leader) mbar_expect_tx(&b_full[s], kBStageFull);
                        if (CG == 1) {
                            bulk_load_1d(b_buf + s * kBStage, wsrc + (size_t)st * kBStageFull, kBStageFull, &b_full[s]);
                        } else {
                            // weight image as rows of 128 B, 128 rows per stage.  X stage (one N = 128 operand): this CTA's 64
                            // rows are contiguous; Y stage (two N = 64 operands): 32 rows of each tile.
                            const uint32_t bar = mapa_rank(smem_u32(&b_full[s]), 0);
                            const int row0 = st * 128, xy = st & 1;
                            const int ra = row0 + (xy ? (int)rank * 32 : (int)rank * 64), rb = xy ? row0 + 64 + (int)rank * 32 : ra + 32;
                            tma_load_2d_cg2(b_buf + s * kBStage, &tmap_w, bar, 0, ra);
                            tma_load_2d_cg2(b_buf + s * kBStage + kBStage / 2, &tmap_w, bar, 0, rb);
                        }
                    }
                    __syncwarp();
                }
    } else if (warp == 1 || warp == 3) {
      if (leader) {
        // ================= MMA issuers: warp 1 owns M tile 0 (hour plane t0), warp 3 owns M tile 1 (t0 + 1) =================
        // The whole warp walks the schedule; one elected lane issues.  Both issuers consume every weight stage.
        const int m = warp == 3 ? 1 : 0;
        constexpr uint32_t kF = HalfOps<HT>::kFmt;
        constexpr uint32_t kM = 128 * CG;
        constexpr uint32_t idesc128 = (1u << 4) | (kF << 7) | (kF << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
        constexpr uint32_t idesc64 = (1u << 4) | (kF << 7) | (kF << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
        const uint32_t a_base = smem_u32(a_reg), b_base = smem_u32(b_buf);
        uint32_t bi = 0, acc_it = 0, li0 = 0;        // li0 = plane-load index of plane 0 of the current unit
        uint32_t b_ready = 0;
        for (int item = grp0; item < n_items; item += grp_stride) {
            const int sup_lo = (item % items_per_grp) * sup_per, sup_hi = sup_lo + sup_per;
            const int p_lo = 2 * sup_lo - 1 < 0 ? 0 : 2 * sup_lo - 1, p_hi = 2 * sup_hi > T - 1 ? T - 1 : 2 * sup_hi;
            int ready = p_lo - 1;                    // highest plane of this item known to have landed
            for (int sup = sup_lo; sup < sup_hi; ++sup)
                for (int pass = 0; pass < 4; ++pass, ++acc_it) {
                    const int pt = pass >> 1, ph = pass & 1;
                    const int w = 2 * sup + pt;      // window: planes w-1, w, w+1
                    const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
                    if (CG == 1) mbar_wait(&acc_empty[as * 2 + m], aph ^ 1);
                    else mbar_wait_cluster(&acc_empty[as * 2 + m], aph ^ 1);
                    tc_fence_after();
                    const uint32_t d_acc = tmem_base + as * kAccCols + m * 128;
                    uint32_t started = 0;
#pragma unroll 1
                    for (int g = 0; g < 8; ++g) {    // g = (at, ah, chunk)
                        const int at = g >> 2, ah = (g >> 1) & 1, c = g & 1;
                        const int dh = ph - 1 + ah;
                        // activation view of this M tile for (dt, dh, chunk); the plane may lie outside [0,T) (zero padding in t)
                        const int plane = 2 * sup + m + pt - 1 + at;
                        const bool a_ok = plane >= 0 && plane < T;
                        const uint32_t l = li0 + (uint32_t)(a_ok ? plane - p_lo : 0);     // ring position: planes of an item load in order
                        if (a_ok) {
                            while (ready < plane) {
                                ++ready;
                                const uint32_t lr = li0 + (uint32_t)(ready - p_lo);
                                mbar_wait(&a_full[lr % kSlots], (lr / kSlots) & 1);
                            }
                        }
                        const uint32_t a_view = a_base + ((l % kSlots) * kChunks + c) * kBufBytes + ((1 + dh) * kLine + 1) * kRow;
                        // ---- stage X: dw = 0, both w-phases in one N = 128 MMA
                        {
                            const uint32_t s = bi % kBStages, bph = (bi / kBStages) & 1;
                            if (!b_ready) mbar_wait(&b_full[s], bph);
                            ++bi;
                            b_ready = mbar_test(&b_full[bi % kBStages], (bi / kBStages) & 1);
                            tc_fence_after();
                            const uint32_t b_addr = b_base + s * kBStage;
                            if (elect_one()) {
                                if (a_ok) {
                                    const uint64_t ad = make_sdesc_sbo(a_view, kSboA), bd = make_sdesc(b_addr);
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        tc_mma_g<CG>(d_acc, ad + 2 * k, bd + 2 * k, idesc128, started | (k ? 1u : 0u));
                                }
                                tc_commit_g<CG>(&b_empty[s]);
                            }
                            if (a_ok) started = 1u;
                            __syncwarp();
                        }
                        // ---- stage Y: dw = -1 feeds pw = 0 (tile 0), dw = +1 feeds pw = 1 (tile 1)
                        {
                            const uint32_t s = bi % kBStages, bph = (bi / kBStages) & 1;
                            if (!b_ready) mbar_wait(&b_full[s], bph);
                            ++bi;
                            b_ready = mbar_test(&b_full[bi % kBStages], (bi / kBStages) & 1);
                            tc_fence_after();
                            const uint32_t b_addr = b_base + s * kBStage;
                            if (elect_one()) {
                                if (a_ok) {
                                    const uint64_t al = make_sdesc_sbo(a_view - kRow, kSboA), ar = make_sdesc_sbo(a_view + kRow, kSboA);
                                    const uint64_t b0d = make_sdesc(b_addr), b1d = make_sdesc(b_addr + kBStage / 2);
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        tc_mma_g<CG>(d_acc, al + 2 * k, b0d + 2 * k, idesc64, 1u);
#pragma unroll
                                    for (int k = 0; k < 4; ++k)
                                        tc_mma_g<CG>(d_acc + 64, ar + 2 * k, b1d + 2 * k, idesc64, 1u);
                                }
                                tc_commit_g<CG>(&b_empty[s]);
                            }
                            __syncwarp();
                        }
                        // plane w-1 is read for the last time by the at = 0 half of the ph = 1 pass of window w
                        if (g == 3 && ph == 1 && w - 1 >= p_lo) {
                            const uint32_t lw = li0 + (uint32_t)(w - 1 - p_lo);
                            if (elect_one()) tc_commit_g<CG>(&a_empty[lw % kSlots]);
                            __syncwarp();
                        }
                    }
                    if (elect_one()) {
                        if (sup == sup_hi - 1 && pass == 3)               // last pass of the item: its remaining planes are done too
                            for (int pl = 2 * sup_hi - 1; pl <= p_hi; ++pl) tc_commit_g<CG>(&a_empty[(li0 + (uint32_t)(pl - p_lo)) % kSlots]);
                        tc_commit_g<CG>(&acc_full[as * 2 + m]);
                    }
                    __syncwarp();
                }
            li0 += (uint32_t)(p_hi - p_lo + 1);
        }
      }
    } else {
        // ================= epilogue =================
        // Per accumulator (M tile m, w-phase pw; thread = position = TMEM lane): read the 64 channels once, bias +
        // PixelNorm + LeakyReLU in f32, round to 16 bit.  Fused mode: the 16-bit activations go back into the drained
        // accumulator's own columns 0..31 (two per column) and the output conv is one more tensor-core product with A
        // from TMEM:  P[pos][tap] = sum_c y[pos][c] * w4[tap][c]  -> columns 32..63; while it completes the group
        // already works on its second accumulator (software pipelined), then stores P (f32, 32 per position).
        const int q = warp & 3;                 // TMEM lane quarter this warp may touch
        const int m = warp >= 8 ? 1 : 0;        // epilogue warpgroup = M tile
        const int r = q * 32 + lane;            // accumulator row: r = (h*2 + sample)*8 + w
        const int h = r >> 4, bl = (r >> 3) & 1, wq = r & 7;
        HT* out = reinterpret_cast<HT*>(args.out);
        const int H2 = 16, W2 = 16, T2 = 2 * T;
        constexpr uint32_t kF = HalfOps<HT>::kFmt;
        constexpr uint32_t idesc_p = (1u << 4) | (kF << 7) | (kF << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
        const uint64_t wd = make_sdesc(smem_u32(w4_tile));
        uint32_t acc_it = 0, p_it = 0;
        const int et = threadIdx.x - 128;       // 0..255 over both epilogue groups
        const float fx_scale = logit_mode ? reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(args.w4tile) + kW4Tile)[0] : 0.f;
        bool bad = false;
        for (int item = grp0; item < n_items; item += grp_stride) {
            const int grp = item / items_per_grp, sup_lo = (item % items_per_grp) * sup_per, sup_hi = sup_lo + sup_per;
            const int unit = grp * CG + (int)rank;
            const int b = unit * 2 + bl;
            const bool valid = b < args.B;
            for (int sup = sup_lo; sup < sup_hi; ++sup)
                for (int pass = 0; pass < 4; ++pass, ++acc_it) {
                    const int pt = pass >> 1, ph = pass & 1;
                    const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
                    mbar_wait(&acc_full[as * 2 + m], aph);
                    tc_fence_after();
                    const size_t o_row = (((size_t)b * T2 + (2 * (2 * sup + m) + pt)) * H2 + (2 * h + ph)) * W2 + 2 * wq;
                    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + as * kAccCols + m * 128;
                    {
#pragma unroll 1
                        for (int pw = 0; pw < 2; ++pw) {
                            const uint32_t taddr = lane_base + pw * 64;
                            uint32_t v0[32], v1[32];
                            tc_ld32_nowait(taddr, v0);
                            tc_ld32_nowait(taddr + 32, v1);
                            tc_ld_wait();
                            float ss = 0.f;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float x0 = __uint_as_float(v0[j]) + s_bias[j], x1 = __uint_as_float(v1[j]) + s_bias[32 + j];
                                v0[j] = __float_as_uint(x0); v1[j] = __float_as_uint(x1);
                                ss = fmaf(x0, x0, ss); ss = fmaf(x1, x1, ss);
                            }
                            const float inv = 1.0f / sqrtf(ss * (1.0f / 64) + 1.0e-8f);
                            bad |= !(ss < INFINITY);                       // Inf / NaN anywhere upstream ends up in ss
                            uint32_t pk[32];
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                float a0 = __uint_as_float(v0[2 * j]) * inv, a1 = __uint_as_float(v0[2 * j + 1]) * inv;
                                float c0 = __uint_as_float(v1[2 * j]) * inv, c1 = __uint_as_float(v1[2 * j + 1]) * inv;
                                a0 = a0 > 0.f ? a0 : 0.2f * a0; a1 = a1 > 0.f ? a1 : 0.2f * a1;
                                c0 = c0 > 0.f ? c0 : 0.2f * c0; c1 = c1 > 0.f ? c1 : 0.2f * c1;
                                pk[j] = HalfOps<HT>::pack(a0, a1);
                                pk[16 + j] = HalfOps<HT>::pack(c0, c1);
                            }
                            if (fuse) {
                                tc_st32(taddr, pk);                        // y as the A operand of the tap product
                                tc_fence_before();
                                asm volatile("bar.sync %0, 128;" ::"r"(1 + m) : "memory");
                                if (q == 0) {
                                    if (CG > 1) {
                                        // both CTAs' y tiles must be in TMEM before the leader issues the pair's product
                                        if (elect_one()) mbar_arrive_cluster(mapa_rank(smem_u32(&y_ready[m * 2 + pw]), 0));
                                        __syncwarp();
                                        if (leader) mbar_wait_cluster(&y_ready[m * 2 + pw], p_it & 1);
                                    }
                                    tc_fence_after();
                                    if (leader && elect_one()) {
                                        const uint32_t a_t = tmem_base + as * kAccCols + m * 128 + pw * 64;   // lane 0
#pragma unroll
                                        for (int k = 0; k < 4; ++k) tc_mma_ts_g<CG>(a_t + 32, a_t + 8 * k, wd + 2 * k, idesc_p, k > 0 ? 1u : 0u);
                                        tc_commit_g<CG>(&p_full[m * 2 + pw]);
                                    }
                                    __syncwarp();
                                }
                            } else if (valid) {
                                uint4* dst = reinterpret_cast<uint4*>(out + (o_row + pw) * 64);
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                            }
                        }
                        if (fuse) {
#pragma unroll 1
                            for (int pw = 0; pw < 2; ++pw) {
                                mbar_wait(&p_full[m * 2 + pw], p_it & 1);
                                tc_fence_after();
                                uint32_t pv[32];
                                tc_ld32(lane_base + pw * 64 + 32, pv);
                                if (logit_mode) {
                                    // out[q] = sum_k P[q + (k - 1)][k]: this position feeds the 27 logits q = pos + 1 - k.
                                    // Fixed-point integer adds: order independent, so results are bit-reproducible.
                                    const int tq = 2 * (2 * sup + m) + pt, hq = 2 * h + ph, wq2 = 2 * wq + pw;
                                    const uint32_t base = smem_u32(ring) + (uint32_t)(hq * kRingPitch + wq2 * 2 + bl) * 4u;
#pragma unroll
                                    for (int kt = 0; kt < 3; ++kt) {
                                        const int to = tq + 1 - kt;
                                        if (to < 0 || to >= T2) continue;            // warp-uniform
                                        const uint32_t bt = base + (uint32_t)((to & (kRingSlots - 1)) * kRingPlane) * 4u;
#pragma unroll
                                        for (int kh = 0; kh < 3; ++kh) {
                                            const bool ok_h = (unsigned)(hq + 1 - kh) < 16u;
#pragma unroll
                                            for (int kw = 0; kw < 3; ++kw) {
                                                const bool ok_w = (unsigned)(wq2 + 1 - kw) < 16u;
                                                const int v = __float2int_rn(__uint_as_float(pv[(kt * 3 + kh) * 3 + kw]) * fx_scale);
                                                if (ok_h && ok_w)
                                                    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(bt + (uint32_t)(((1 - kh) * kRingPitch + (1 - kw) * 2) * 4)), "r"(v) : "memory");
                                            }
                                        }
                                    }
                                } else if (valid) {
                                    uint4* dst = reinterpret_cast<uint4*>(args.p_out + (o_row + pw) * 32);
#pragma unroll
                                    for (int j = 0; j < 8; ++j)
                                        dst[j] = make_uint4(pv[4 * j], pv[4 * j + 1], pv[4 * j + 2], pv[4 * j + 3]);
                                }
                            }
                        }
                    }
                    ++p_it;
                    tc_fence_before();
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + m) : "memory");       // the whole group has drained the accumulator
                    if ((threadIdx.x & 127) == 0) {
                        if (CG == 1) mbar_arrive(&acc_empty[as * 2 + m]);
                        else mbar_arrive_cluster(mapa_rank(smem_u32(&acc_empty[as * 2 + m]), 0));
                    }
                    if (logit_mode && pass == 3) {
                        // End of a super-tile (low-res planes 2sup, 2sup+1 = hours 4sup..4sup+3): hours up to 4sup+2 have
                        // all their contributions.  Both epilogue groups flush them (raw fixed point; the softmax kernel
                        // finishes in place) and clear the ring slots for reuse.
                        asm volatile("bar.sync 3, 256;" ::: "memory");
                        const int h_lo = sup == 0 ? 0 : 4 * sup - 1;
                        const int h_hi = sup == sup_hi - 1 ? (4 * sup + 4 > T2 - 1 ? T2 - 1 : 4 * sup + 4) : 4 * sup + 2;
                        const int n = (h_hi - h_lo + 1) * 512;
                        for (int i = et; i < n; i += 256) {
                            const int hour = h_lo + (i >> 9), s2 = (i >> 8) & 1, pix = i & 255;
                            int* rp = ring + (hour & (kRingSlots - 1)) * kRingPlane + (pix >> 4) * kRingPitch + (pix & 15) * 2 + s2;
                            const int v = *rp;
                            *rp = 0;
                            const int bb = unit * 2 + s2;
                            if (bb < args.B) {
                                int* dst = args.logit_out + ((size_t)bb * T2 + hour) * 256 + pix;
                                if (args.sup_split) atomicAdd(dst, v); else *dst = v;      // split items share their border hours
                            }
                        }
                        asm volatile("bar.sync 3, 256;" ::: "memory");
                    }
                }
        }
        if (bad && args.nonfinite) atomicOr(args.nonfinite, 1);
    }

    tc_fence_before();
    __syncthreads();
    if (CG > 1) cluster_sync_all();              // no CTA leaves while the pair's MMAs / multicasts may still touch it
    if (warp == 1) {
        tc_fence_after();
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

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

template <typename HT>
int launch_planes(const void* x, const void* wpack, const float* bias, void* y, const void* w4tile, float* p_out,
                  int* logit_out, int* nonfinite, int B, int T, int sm_count, cudaStream_t st) {
    constexpr int H = 8, W = 8, Cin = 128;
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { rdg_set_error("cuTensorMapEncodeTiled entry point not available"); return RDG_TC_E_DRIVER; }
    if (T < 2 || (T & 1)) { rdg_set_error("tc planes: T must be even"); return RDG_TC_E_SHAPE; }
    TcConvArgs a{};
    a.B = B; a.T = T; a.H = H; a.W = W; a.Cin = Cin;
    a.wpack = wpack; a.bias = bias; a.out = y; a.w4tile = w4tile; a.p_out = p_out; a.logit_out = logit_out; a.nonfinite = nonfinite;

    // tensor dims ordered (C, W, B, H, T) so that the box lands as [h'][sample][w'][64 ch] rows
    CUtensorMap tmap;
    cuuint64_t gdim[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)B, (cuuint64_t)H, (cuuint64_t)T};
    cuuint64_t gstr[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)T * H * W * Cin * 2, (cuuint64_t)W * Cin * 2,
                          (cuuint64_t)H * W * Cin * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)kWp, 2, 10, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUtensorMapDataType dt = HalfOps<HT>::kFmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = enc(&tmap, dt, 5, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { rdg_set_error("cuTensorMapEncodeTiled (planes) failed: %d", (int)r); return RDG_TC_E_DRIVER; }

    // folded weight image as rows of 128 B (64 stages x 128 rows), boxes of 32 rows: the 2-CTA kernel's weight loads
    CUtensorMap tmap_w;
    {
        cuuint64_t wdim[2] = {64, 64 * 128};
        cuuint64_t wstr[1] = {128};
        cuuint32_t wbox[2] = {64, 32};
        cuuint32_t wes[2] = {1, 1};
        r = enc(&tmap_w, dt, 2, const_cast<void*>(wpack), wdim, wstr, wbox, wes, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { rdg_set_error("cuTensorMapEncodeTiled (planes weights) failed: %d", (int)r); return RDG_TC_E_DRIVER; }
    }
    const int n_units = (B + 1) / 2;
    // small launches: one work item per (group of units, super-tile) instead of per group (latency of a few samples)
    a.sup_split = ((n_units + 1) / 2) * 2 <= sm_count / 2 ? 1 : 0;
    if (a.sup_split && logit_out) RDG_CUDA(cudaMemsetAsync(logit_out, 0, (size_t)B * 2 * T * 256 * sizeof(int), st));
    // CTA pairs (cta_group::2) halve the weight-operand reads and weight writes per SM; RDG_PLANES_CG=1: one CTA per unit
    static const int cg = getenv("RDG_PLANES_CG") ? atoi(getenv("RDG_PLANES_CG")) : 2;
    if (cg == 2 && n_units >= 2) {
        auto kern = tc_upconv64_planes_kernel<HT, 2>;
        static bool attr_set = false;
        if (!attr_set) {
            RDG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
            attr_set = true;
        }
        const int n_groups = (n_units + 1) / 2, max_clusters = sm_count / 2;
        const int n_items = n_groups * (a.sup_split ? T / 2 : 1);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(2 * (n_items < max_clusters ? n_items : max_clusters));
        cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        RDG_CUDA(cudaLaunchKernelEx(&cfg, kern, tmap, tmap_w, a));
        return 0;
    }
    auto kern = tc_upconv64_planes_kernel<HT, 1>;
    static bool attr_set = false;
    if (!attr_set) {
        RDG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_set = true;
    }
    const int n_items1 = n_units * (a.sup_split ? T / 2 : 1);
    const int grid = n_items1 < sm_count ? n_items1 : sm_count;
    kern<<<grid, kThreads, kSmem, st>>>(tmap, tmap_w, a);
    RDG_LAUNCH_CHECK();
    return 0;
}

// Folded weights of the 128 -> 64 layer in the stage order of tc_upconv64_planes_kernel:
// [pt][ph][at][ah][chunk][X|Y][tile 2][64 cout rows x 64 k], rows 128B-swizzled.  X tiles: (pw0,aw1), (pw1,aw0);
// Y tiles: (pw0,aw0), (pw1,aw1).  Fold rule as in pack_folded_kernel: sums in f32, rounded once.
template <typename HT>
__global__ void pack_folded_planes_kernel(const float* __restrict__ k, HT* __restrict__ dst) {
    constexpr int Cin = 128, Cout = 64;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 64 * 8192) return;
    const int sidx = idx >> 13, el = idx & 8191;
    const int j = el >> 12, n = (el >> 6) & 63, pos = el & 63;
    const int xy = sidx & 1, c = (sidx >> 1) & 1, ah = (sidx >> 2) & 1, at = (sidx >> 3) & 1, ph = (sidx >> 4) & 1, pt = (sidx >> 5) & 1;
    const int pw = j, aw = xy == 0 ? 1 - j : j;
    const int ci = c * 64 + (((pos >> 3) ^ (n & 7)) << 3) + (pos & 7);
    auto lo = [](int phs, int tp) { return phs == 0 ? (tp == 0 ? 0 : 1) : (tp == 0 ? 0 : 2); };
    auto hi = [](int phs, int tp) { return phs == 0 ? (tp == 0 ? 0 : 2) : (tp == 0 ? 1 : 2); };
    float sum = 0.f;
    for (int kt = lo(pt, at); kt <= hi(pt, at); ++kt)
        for (int kh = lo(ph, ah); kh <= hi(ph, ah); ++kh)
            for (int kw = lo(pw, aw); kw <= hi(pw, aw); ++kw)
                sum += k[((size_t)((kt * 3 + kh) * 3 + kw) * Cin + ci) * Cout + n];
    dst[idx] = HalfOps<HT>::from_float(sum);
}

}  // namespace

int tc_upconv64_planes(int half_kind, const void* x, const void* wpack, const float* bias, void* y, const void* w4tile,
                       float* p_out, int* logit_out, int* nonfinite, int B, int T, int sm_count, cudaStream_t st) {
    if (B <= 0) return 0;
    if (half_kind == RDG_HALF_BF16)
        return launch_planes<__nv_bfloat16>(x, wpack, bias, y, w4tile, p_out, logit_out, nonfinite, B, T, sm_count, st);
    return launch_planes<__half>(x, wpack, bias, y, w4tile, p_out, logit_out, nonfinite, B, T, sm_count, st);
}

int pack_folded_weights_planes(int half_kind, const float* k, void* dst, cudaStream_t st) {
    if (half_kind == RDG_HALF_BF16) pack_folded_planes_kernel<__nv_bfloat16><<<64 * 8192 / 256, 256, 0, st>>>(k, (__nv_bfloat16*)dst);
    else pack_folded_planes_kernel<__half><<<64 * 8192 / 256, 256, 0, st>>>(k, (__half*)dst);
    RDG_LAUNCH_CHECK();
    return 0;
}
