// FP32 SIMT 3-D convolution primitives (forward, backward-data, backward-filter) over
// channels-last tensors, as implicit GEMMs with 64x64x16 shared-memory tiles and 4x4
// register micro-tiles.  These are the FP32-mode / training building blocks; the BF16/FP16
// generator forward uses the tcgen05 kernels in gen_tc.cu instead.
//
// Semantics follow Keras Conv3D as used by the reference (cross-correlation, kernel layout
// (kt,kh,kw,Cin,Cout), TF 'same'/'valid' padding expressed as pad-before + output size):
// gan_train_cwgangp_pixelnorm.py:286-301 (critic, stride 2) and :331-345 (generator, stride 1
// after UpSampling3D, which is fused into the gather here via ConvGeom::up).
#include "rdg_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct PosDec { int b, t, h, w; };

__device__ __forceinline__ PosDec decode_pos(long long m, int T, int H, int W) {
    PosDec p;
    p.w = (int)(m % W); m /= W;
    p.h = (int)(m % H); m /= H;
    p.t = (int)(m % T); p.b = (int)(m / T);
    return p;
}

__device__ __forceinline__ void fma_tile(float (&acc)[4][4], const float (*As)[BM + 4], const float (*Bs)[BN + 4],
                                         int ty, int tx) {
#pragma unroll
    for (int k = 0; k < BK; ++k) {
        float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// MODE 0: forward.   M = output positions, N = Co, K = taps*Ci
// MODE 1: bwd-data.  M = logical input positions, N = Ci, K = taps*Co
// Split-K (gridDim.z > 1): slice z reduces a contiguous range of the (tap, channel-chunk) steps and writes its raw
// partial tile to part[z][M][N]; splitk_epilogue_kernel then sums the slices in fixed order (deterministic) and applies
// the epilogue.  Used when the M x N tile grid alone cannot fill the GPU (the critic's deep layers: M = 64 .. 3072 rows
// at batch 32 with K up to 6912), where a few CTAs walking hundreds of dependent load->sync->FMA steps are latency-bound.
template <int MODE>
__global__ void __launch_bounds__(NT) conv_gemm_kernel(const float* __restrict__ src, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ dst,
                                                       ConvGeom g, int act, const float* __restrict__ mask,
                                                       float mask_scale, float* __restrict__ pre, float* __restrict__ part) {
    // double-buffered tiles: while the FMAs of step s run, the global loads of step s+1 are in flight in registers
    // (one __syncthreads per step; small grids are bound by the load latency of the step chain, not by FLOPs)
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int upf = g.up ? 2 : 1;
    const int LT = g.Ti * upf, LH = g.Hi * upf, LW = g.Wi * upf;   // logical input dims
    const int MT = MODE == 0 ? g.To : LT, MH = MODE == 0 ? g.Ho : LH, MW = MODE == 0 ? g.Wo : LW;
    const long long M = (long long)g.B * MT * MH * MW;
    const int N = MODE == 0 ? g.Co : g.Ci;
    const int KC = MODE == 0 ? g.Ci : g.Co;                        // reduction channels per tap
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    const int a_m = tid >> 2, a_k = (tid & 3) * 4;
    const long long am = m0 + a_m;
    const bool am_ok = am < M;
    const PosDec ap = decode_pos(am_ok ? am : 0, MT, MH, MW);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int ntaps = g.KT * g.KH * g.KW;
    const int nck = (KC + BK - 1) / BK;                            // channel chunks per tap
    const int total_steps = ntaps * nck;
    const int per_slice = (total_steps + (int)gridDim.z - 1) / (int)gridDim.z;
    const int step_beg = (int)blockIdx.z * per_slice;
    const int step_end = step_beg + per_slice < total_steps ? step_beg + per_slice : total_steps;

    // gather state of the current tap for this thread's A row
    int cur_tap = -1;
    bool ok = false;
    long long base = 0;
    auto set_tap = [&](int tap) {
        cur_tap = tap;
        const int kw_ = tap % g.KW, kh_ = (tap / g.KW) % g.KH, kt_ = tap / (g.KW * g.KH);
        ok = am_ok;
        if (MODE == 0) {
            int lt = ap.t * g.stride + kt_ - g.pt, lh = ap.h * g.stride + kh_ - g.ph, lw = ap.w * g.stride + kw_ - g.pw;
            ok = ok && lt >= 0 && lt < LT && lh >= 0 && lh < LH && lw >= 0 && lw < LW;
            if (g.up) { lt >>= 1; lh >>= 1; lw >>= 1; }
            base = ((((long long)ap.b * g.Ti + lt) * g.Hi + lh) * g.Wi + lw) * g.Ci;
        } else {
            int nt = ap.t + g.pt - kt_, nh = ap.h + g.ph - kh_, nw = ap.w + g.pw - kw_;
            ok = ok && nt >= 0 && nh >= 0 && nw >= 0 && (nt % g.stride) == 0 && (nh % g.stride) == 0 &&
                 (nw % g.stride) == 0;
            nt /= g.stride; nh /= g.stride; nw /= g.stride;
            ok = ok && nt < g.To && nh < g.Ho && nw < g.Wo;
            base = ((((long long)ap.b * g.To + nt) * g.Ho + nh) * g.Wo + nw) * g.Co;
        }
    };
    float av[4], bv[4];
    // global -> registers for reduction step `step` (A: 64 positions x 16 channels, B: 16 channels x 64 outputs)
    auto load_step = [&](int step) {
        const int tap = step / nck, c0 = (step - tap * nck) * BK;
        if (tap != cur_tap) set_tap(tap);
        av[0] = av[1] = av[2] = av[3] = 0.f;
        if (ok) {
            const float* p = src + base + c0 + a_k;
            if (c0 + a_k + 3 < KC && ((KC & 3) == 0)) {
                const float4 v = *reinterpret_cast<const float4*>(p);
                av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) if (c0 + a_k + j < KC) av[j] = p[j];
            }
        }
        if (MODE == 0) {
            const int bk = tid >> 4, bn = (tid & 15) * 4, ci = c0 + bk;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + bn + j;
                bv[j] = (ci < g.Ci && n < g.Co) ? w[((long long)tap * g.Ci + ci) * g.Co + n] : 0.f;
            }
        } else {
            const int bk = tid & 15, co = c0 + bk;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ci = n0 + (tid >> 4) + j * 16;
                bv[j] = (co < g.Co && ci < g.Ci) ? w[((long long)tap * g.Ci + ci) * g.Co + co] : 0.f;
            }
        }
    };
    auto store_step = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 4; ++j) As[buf][a_k + j][a_m] = av[j];
        if (MODE == 0) {
            const int bk = tid >> 4, bn = (tid & 15) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) Bs[buf][bk][bn + j] = bv[j];
        } else {
            const int bk = tid & 15;
#pragma unroll
            for (int j = 0; j < 4; ++j) Bs[buf][bk][(tid >> 4) + j * 16] = bv[j];
        }
    };
    if (step_beg < step_end) {
        load_step(step_beg);
        store_step(0);
        __syncthreads();
        for (int step = step_beg; step < step_end; ++step) {
            const int buf = (step - step_beg) & 1;
            const bool more = step + 1 < step_end;
            if (more) load_step(step + 1);                 // in flight during the FMAs below
            fma_tile(acc, As[buf], Bs[buf], ty, tx);
            if (more) store_step(buf ^ 1);                 // the other buffer was last read before the previous barrier
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (part) { part[((long long)blockIdx.z * M + m) * N + n] = v; continue; }
            if (MODE == 0) {
                if (bias) v += bias[n];
                if (pre) pre[m * N + n] = v;            // pre-activation, kept for the backward pass
                if (act == ACT_LRELU) v = v > 0.f ? v : 0.2f * v;
                if (mask) v *= mask[m * N + n] * mask_scale;
            } else if (act == ACT_ACCUM) {
                v += dst[m * N + n];                    // backward-data of one phase of a folded conv: sum over the phases
            }
            dst[m * N + n] = v;
        }
    }
}

// dst = epilogue(sum_z part[z]); same epilogue as conv_gemm_kernel (bias, pre-activation copy, LeakyReLU, dropout mask)
__global__ void splitk_epilogue_kernel(const float* __restrict__ part, int nslice, long long MN, int N,
                                       const float* __restrict__ bias, float* __restrict__ dst, int act,
                                       const float* __restrict__ mask, float mask_scale, float* __restrict__ pre) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= MN) return;
    float v = 0.f;
    for (int z = 0; z < nslice; ++z) v += part[(long long)z * MN + i];
    if (bias) v += bias[i % N];
    if (pre) pre[i] = v;
    if (act == ACT_ACCUM) v += dst[i];
    if (act == ACT_LRELU) v = v > 0.f ? v : 0.2f * v;
    if (mask) v *= mask[i] * mask_scale;
    dst[i] = v;
}

// dW[tap][ci][co] += sum_m x[gather(m,tap)][ci] * dy[m][co]; grid (ci tiles, co tiles, taps*ksplit)
__global__ void __launch_bounds__(NT) conv_bwd_filter_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                             float* __restrict__ dw, ConvGeom g, int ksplit) {
    __shared__ __align__(16) float As[BK][BM + 4];   // [pos][ci]
    __shared__ __align__(16) float Bs[BK][BN + 4];   // [pos][co]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int upf = g.up ? 2 : 1;
    const int LT = g.Ti * upf, LH = g.Hi * upf, LW = g.Wi * upf;
    const long long M = (long long)g.B * g.To * g.Ho * g.Wo;
    const int tap = blockIdx.z / ksplit, ks = blockIdx.z % ksplit;
    const int kw_ = tap % g.KW, kh_ = (tap / g.KW) % g.KH, kt_ = tap / (g.KW * g.KH);
    const int ci0 = blockIdx.x * BM, co0 = blockIdx.y * BN;
    const long long per = ((M + ksplit - 1) / ksplit + BK - 1) / BK * BK;
    const long long mbeg = ks * per, mend = mbeg + per < M ? mbeg + per : M;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int lp = tid >> 4;            // position within chunk (0..15)
    const int lc = (tid & 15) * 4;      // channel quad
    for (long long mc = mbeg; mc < mend; mc += BK) {
        long long m = mc + lp;
        bool okm = m < mend;
        PosDec p = decode_pos(okm ? m : 0, g.To, g.Ho, g.Wo);
        int lt = p.t * g.stride + kt_ - g.pt, lh = p.h * g.stride + kh_ - g.ph, lw = p.w * g.stride + kw_ - g.pw;
        bool okx = okm && lt >= 0 && lt < LT && lh >= 0 && lh < LH && lw >= 0 && lw < LW;
        if (g.up) { lt >>= 1; lh >>= 1; lw >>= 1; }
        const float* xp = x + ((((long long)p.b * g.Ti + lt) * g.Hi + lh) * g.Wi + lw) * g.Ci;
        const float* yp = dy + m * g.Co;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int ci = ci0 + lc + j, co = co0 + lc + j;
            As[lp][lc + j] = (okx && ci < g.Ci) ? xp[ci] : 0.f;
            Bs[lp][lc + j] = (okm && co < g.Co) ? yp[co] : 0.f;
        }
        __syncthreads();
        fma_tile(acc, As, Bs, ty, tx);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int ci = ci0 + ty * 4 + i;
        if (ci >= g.Ci) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int co = co0 + tx * 4 + j;
            if (co >= g.Co) continue;
            atomicAdd(&dw[((long long)tap * g.Ci + ci) * g.Co + co], acc[i][j]);
        }
    }
}

// ---- few-input-channel layers (the critic's first conv: Ci = 1 + ncond, gan_train_cwgangp_pixelnorm.py:286-287) ----
// The 64x64-tile kernels waste 62/64 of a tile on them.  dx[m][ci] = sum_{valid taps} sum_co dy[n(m,tap)][co] * w[tap][ci][co]:
// one thread per (input position, ci), weights in shared memory, dy rows read as float4.
__global__ void __launch_bounds__(256) conv_bwd_data_smallci_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                    float* __restrict__ dx, ConvGeom g, int ci_n) {
    extern __shared__ __align__(16) float ws[];              // [tap][ci][co]
    const int ntaps = g.KT * g.KH * g.KW;
    for (int i = threadIdx.x; i < ntaps * g.Ci * g.Co; i += blockDim.x) ws[i] = w[i];
    __syncthreads();
    const long long M = (long long)g.B * g.Ti * g.Hi * g.Wi;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * ci_n) return;          // ci_n < Ci: only the first ci_n input channels are wanted (the others are left untouched)
    const int ci = (int)(idx % ci_n);
    const PosDec p = decode_pos(idx / ci_n, g.Ti, g.Hi, g.Wi);
    float acc = 0.f;
    for (int kt = 0; kt < g.KT; ++kt) {
        int nt = p.t + g.pt - kt;
        if (nt < 0 || nt % g.stride) continue;
        nt /= g.stride;
        if (nt >= g.To) continue;
        for (int kh = 0; kh < g.KH; ++kh) {
            int nh = p.h + g.ph - kh;
            if (nh < 0 || nh % g.stride) continue;
            nh /= g.stride;
            if (nh >= g.Ho) continue;
            for (int kw = 0; kw < g.KW; ++kw) {
                int nw = p.w + g.pw - kw;
                if (nw < 0 || nw % g.stride) continue;
                nw /= g.stride;
                if (nw >= g.Wo) continue;
                const float4* yp = reinterpret_cast<const float4*>(dy + ((((long long)p.b * g.To + nt) * g.Ho + nh) * g.Wo + nw) * g.Co);
                const float4* wp = reinterpret_cast<const float4*>(ws + ((size_t)((kt * g.KH + kh) * g.KW + kw) * g.Ci + ci) * g.Co);
                for (int c = 0; c < g.Co / 4; ++c) {
                    const float4 a = yp[c], b = wp[c];
                    acc = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
                }
            }
        }
    }
    dx[(idx / ci_n) * g.Ci + ci] = acc;
}

// dW[tap][ci][co] += sum_m x[gather(m,tap)][ci] * dy[m][co] when Ci * Co <= 256 (few input OR few output channels: the
// critic's first conv, the generator's output conv): block = (tap, m-slice), thread = (ci, co); four positions per
// iteration so that their loads are in flight together (the loop is otherwise one dependent global load per step).
__global__ void __launch_bounds__(512) conv_bwd_filter_small_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                    float* __restrict__ dw, ConvGeom g, int nslice) {
    constexpr int kMaxRows = 2048;
    __shared__ int soff[kMaxRows];           // gather offset of each of the block's rows for this tap (-1: zero padding)
    __shared__ float red[512];
    const int tap = blockIdx.x, slice = blockIdx.y;
    const int kw_ = tap % g.KW, kh_ = (tap / g.KW) % g.KH, kt_ = tap / (g.KW * g.KH);
    const int npair = g.Ci * g.Co, nsub = blockDim.x / npair;      // row subgroups per block, reduced in shared memory
    const int pair = threadIdx.x % npair, sub = threadIdx.x / npair;
    const int co = pair % g.Co, ci = pair / g.Co;
    const long long M = (long long)g.B * g.To * g.Ho * g.Wo;
    const long long per_blk = (M + nslice - 1) / nslice;           // host guarantees per_blk <= kMaxRows
    const long long mbeg = (long long)slice * per_blk, mend = mbeg + per_blk < M ? mbeg + per_blk : M;
    const int rows = mend > mbeg ? (int)(mend - mbeg) : 0;
    // the index arithmetic is the same for all (ci, co) pairs: do it once per row
    for (int r = threadIdx.x; r < rows; r += blockDim.x) {
        const PosDec p = decode_pos(mbeg + r, g.To, g.Ho, g.Wo);
        const int lt = p.t * g.stride + kt_ - g.pt, lh = p.h * g.stride + kh_ - g.ph, lw = p.w * g.stride + kw_ - g.pw;
        const bool ok = lt >= 0 && lt < g.Ti && lh >= 0 && lh < g.Hi && lw >= 0 && lw < g.Wi;
        soff[r] = ok ? (int)(((((long long)p.b * g.Ti + lt) * g.Hi + lh) * g.Wi + lw) * g.Ci) : -1;
    }
    __syncthreads();
    float acc = 0.f;
    if (sub < nsub) {
        const float* dyb = dy + mbeg * g.Co + co;
        constexpr int U = 16;                 // rows in flight per thread: the loop is bound by the latency of its loads
        for (int r = sub; r < rows; r += U * nsub) {
            float xv[U], yv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int rr = r + u * nsub;
                const int off = rr < rows ? soff[rr] : -1;
                xv[u] = off >= 0 ? x[off + ci] : 0.f;
                yv[u] = off >= 0 ? dyb[(size_t)rr * g.Co] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc = fmaf(xv[u], yv[u], acc);
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (sub == 0) {
        for (int k = 1; k < nsub; ++k) acc += red[k * npair + pair];
        atomicAdd(&dw[((long long)tap * g.Ci + ci) * g.Co + co], acc);
    }
}

// y[m][co] = act(sum_{tap,ci} x[gather(m,tap)][ci] * w[tap][ci][co] + bias) for Co <= 4 (the generator's output conv in the
// training path): one thread per (position, co), weights transposed into shared memory as [tap][co][ci].
__global__ void __launch_bounds__(256) conv_fwd_smallco_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float* __restrict__ y,
                                                               ConvGeom g, int act, float* __restrict__ pre) {
    extern __shared__ __align__(16) float ws[];              // [tap][co][ci]
    const int ntaps = g.KT * g.KH * g.KW;
    for (int i = threadIdx.x; i < ntaps * g.Ci * g.Co; i += blockDim.x) {
        const int co = i % g.Co, ci = (i / g.Co) % g.Ci, tap = i / (g.Co * g.Ci);
        ws[((size_t)tap * g.Co + co) * g.Ci + ci] = w[i];
    }
    __syncthreads();
    const long long M = (long long)g.B * g.To * g.Ho * g.Wo;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * g.Co) return;
    const int co = (int)(idx % g.Co);
    const PosDec p = decode_pos(idx / g.Co, g.To, g.Ho, g.Wo);
    float acc0 = 0.f, acc1 = 0.f;
    for (int kt = 0; kt < g.KT; ++kt) {
        const int lt = p.t * g.stride + kt - g.pt;
        if (lt < 0 || lt >= g.Ti) continue;
        for (int kh = 0; kh < g.KH; ++kh) {
            const int lh = p.h * g.stride + kh - g.ph;
            if (lh < 0 || lh >= g.Hi) continue;
            for (int kw = 0; kw < g.KW; ++kw) {
                const int lw = p.w * g.stride + kw - g.pw;
                if (lw < 0 || lw >= g.Wi) continue;
                const float4* xp = reinterpret_cast<const float4*>(x + ((((long long)p.b * g.Ti + lt) * g.Hi + lh) * g.Wi + lw) * g.Ci);
                const float4* wp = reinterpret_cast<const float4*>(ws + ((size_t)((kt * g.KH + kh) * g.KW + kw) * g.Co + co) * g.Ci);
                for (int c = 0; c < g.Ci / 4; c += 2) {
                    const float4 a = xp[c], b = wp[c], a2 = xp[c + 1], b2 = wp[c + 1];
                    acc0 = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc0))));
                    acc1 = fmaf(a2.x, b2.x, fmaf(a2.y, b2.y, fmaf(a2.z, b2.z, fmaf(a2.w, b2.w, acc1))));
                }
            }
        }
    }
    float v = acc0 + acc1;
    if (bias) v += bias[co];
    if (pre) pre[idx] = v;
    if (act == ACT_LRELU) v = v > 0.f ? v : 0.2f * v;
    y[idx] = v;
}

// db[c] += sum_m dy[m][c]: a block sums rows_per_block rows; 8 row groups (warps) x 32 channel lanes, partial sums in double
// (the fake / real halves of a critic batch nearly cancel), reduced over the row groups in shared memory, one float atomic per
// channel and block.
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dy, float* __restrict__ db, long long M, int C,
                                                     long long rows_per_block) {
    __shared__ double red[8][33];
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    const long long r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int c = blockIdx.y * 32 + lane;           // one 32-channel block per blockIdx.y
    double s = 0.0;
    if (c < C)
        for (long long r = r0 + grp; r < r1; r += 8) s += (double)dy[r * C + c];
    red[grp][lane] = s;
    __syncthreads();
    if (grp == 0 && c < C) {
        for (int k = 1; k < 8; ++k) s += red[k][lane];
        atomicAdd(&db[c], (float)s);
    }
}

}  // namespace

// ---- backward-data of a stride-2 convolution by parity classes ----
// dx[p] only receives taps with k = (p + pad) mod 2 on every axis, so a generic tap loop spends 7/8 of its steps on
// zeros.  Rows are enumerated class-major (class q = parity of p + pad per axis, 8 classes); a 64-row tile is
// class-homogeneous and walks only its class's taps (k = q, q + 2, ...): 27 taps -> 8/4/4/2/4/2/2/1.
struct ClsInfo { int tile_begin[9]; };

__global__ void __launch_bounds__(NT) conv_bwd_data_s2_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                              float* __restrict__ dx, ConvGeom g, ClsInfo ci_, float* __restrict__ part) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    int cls = 0;
    while (cls < 7 && (int)blockIdx.x >= ci_.tile_begin[cls + 1]) ++cls;
    const int qt = cls >> 2, qh = (cls >> 1) & 1, qw = cls & 1;
    const int ot = (qt - g.pt) & 1, oh = (qh - g.ph) & 1, ow = (qw - g.pw) & 1;       // first position of the class per axis
    const int ct = (g.Ti - ot + 1) / 2, ch = (g.Hi - oh + 1) / 2, cw = (g.Wi - ow + 1) / 2;   // positions of the class per axis
    const long long rows = (long long)g.B * ct * ch * cw;
    const long long r0 = (long long)((int)blockIdx.x - ci_.tile_begin[cls]) * BM;
    const int n0 = blockIdx.y * BN;
    const int nkt = (g.KT - qt + 1) / 2, nkh = (g.KH - qh + 1) / 2, nkw = (g.KW - qw + 1) / 2;
    const int ntaps = nkt * nkh * nkw;

    const int a_m = tid >> 2, a_k = (tid & 3) * 4;
    const bool am_ok = r0 + a_m < rows;
    const PosDec ap = decode_pos(am_ok ? r0 + a_m : 0, ct, ch, cw);
    const int pt_ = ot + 2 * ap.t, ph_ = oh + 2 * ap.h, pw_ = ow + 2 * ap.w;          // actual input position

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int nck = (g.Co + BK - 1) / BK;
    const int total_steps = ntaps * nck;
    const int per_slice = (total_steps + (int)gridDim.z - 1) / (int)gridDim.z;
    const int step_beg = (int)blockIdx.z * per_slice;
    const int step_end = step_beg + per_slice < total_steps ? step_beg + per_slice : total_steps;
    for (int ti = step_beg / nck; ti * nck < step_end; ++ti) {
        const int kt_ = qt + 2 * (ti / (nkh * nkw)), kh_ = qh + 2 * ((ti / nkw) % nkh), kw_ = qw + 2 * (ti % nkw);
        const int tap = (kt_ * g.KH + kh_) * g.KW + kw_;
        const int nt = (pt_ + g.pt - kt_) / 2, nh = (ph_ + g.ph - kh_) / 2, nw = (pw_ + g.pw - kw_) / 2;
        const bool ok = am_ok && pt_ + g.pt >= kt_ && ph_ + g.ph >= kh_ && pw_ + g.pw >= kw_ && nt < g.To && nh < g.Ho && nw < g.Wo;
        const long long base = ((((long long)ap.b * g.To + nt) * g.Ho + nh) * g.Wo + nw) * g.Co;
        const int ck_beg = ti * nck < step_beg ? step_beg - ti * nck : 0;
        const int ck_end = (ti + 1) * nck > step_end ? step_end - ti * nck : nck;
        for (int c0 = ck_beg * BK; c0 < ck_end * BK; c0 += BK) {
            float av[4] = {0.f, 0.f, 0.f, 0.f};
            if (ok) {
                const float* p = dy + base + c0 + a_k;
                if (c0 + a_k + 3 < g.Co && ((g.Co & 3) == 0)) {
                    float4 v = *reinterpret_cast<const float4*>(p);
                    av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (c0 + a_k + j < g.Co) av[j] = p[j];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) As[a_k + j][a_m] = av[j];
            const int bk = tid & 15, co = c0 + bk;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int nn = (tid >> 4) + j * 16, ci = n0 + nn;
                Bs[bk][nn] = (co < g.Co && ci < g.Ci) ? w[((long long)tap * g.Ci + ci) * g.Co + co] : 0.f;
            }
            __syncthreads();
            fma_tile(acc, As, Bs, ty, tx);
            __syncthreads();
        }
    }
    const long long Mtot = (long long)g.B * g.Ti * g.Hi * g.Wi;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = r0 + ty * 4 + i;
        if (r >= rows) continue;
        const PosDec q = decode_pos(r, ct, ch, cw);
        const long long m = (((long long)q.b * g.Ti + (ot + 2 * q.t)) * g.Hi + (oh + 2 * q.h)) * g.Wi + (ow + 2 * q.w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.Ci) continue;
            if (part) part[((long long)blockIdx.z * Mtot + m) * g.Ci + n] = acc[i][j];
            else dx[m * g.Ci + n] = acc[i][j];
        }
    }
}

// Reduction slices for a tile grid of `ctas` CTAs over `steps` K steps: aim at ~2 CTAs per SM, at least 4 steps per slice.
static int splitk_slices(int ctas, int steps) {
    if (ctas >= 148 || steps < 8) return 1;
    int n = (296 + ctas - 1) / ctas;
    if (n > steps / 4) n = steps / 4;
    return n < 1 ? 1 : n;
}

int simt_conv_fwd(const float* x, const float* w, const float* bias, float* y, const ConvGeom& g, int act,
                  const float* mask, float mask_scale, cudaStream_t st, float* pre) {
    long long M = (long long)g.B * g.To * g.Ho * g.Wo;
    if (M == 0) return 0;
    const int ntaps_ = g.KT * g.KH * g.KW;
    if (!g.up && !mask && g.Co <= 4 && (g.Ci & 7) == 0 && (size_t)ntaps_ * g.Ci * g.Co * 4 <= 48 * 1024) {
        conv_fwd_smallco_kernel<<<ceil_div(M * g.Co, 256), 256, (size_t)ntaps_ * g.Ci * g.Co * 4, st>>>(x, w, bias, y, g, act, pre);
        RDG_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid(ceil_div(M, BM), ceil_div(g.Co, BN));
    // slices chosen for a nominal batch of >= 32 so that a sample's summation order (hence its bits) is the same whether it
    // is computed alone or inside a small batch (tests pin batches through exact linearity in the batch)
    const int nslice = splitk_slices(ceil_div(M / g.B * (g.B < 32 ? 32 : g.B), BM) * grid.y, g.KT * g.KH * g.KW * ceil_div(g.Ci, BK));
    if (nslice > 1) {
        float* part = nullptr;
        RDG_CUDA(cudaMallocAsync(&part, (size_t)nslice * M * g.Co * sizeof(float), st));
        grid.z = nslice;
        conv_gemm_kernel<0><<<grid, NT, 0, st>>>(x, w, nullptr, y, g, ACT_NONE, nullptr, 1.f, nullptr, part);
        RDG_LAUNCH_CHECK();
        splitk_epilogue_kernel<<<ceil_div(M * g.Co, 256), 256, 0, st>>>(part, nslice, M * g.Co, g.Co, bias, y, act, mask, mask_scale, pre);
        RDG_LAUNCH_CHECK();
        RDG_CUDA(cudaFreeAsync(part, st));
        return 0;
    }
    conv_gemm_kernel<0><<<grid, NT, 0, st>>>(x, w, bias, y, g, act, mask, mask_scale, pre, nullptr);
    RDG_LAUNCH_CHECK();
    return 0;
}

int simt_conv_bwd_data(const float* dy, const float* w, float* dx, const ConvGeom& g, cudaStream_t st, int accumulate) {
    int upf = g.up ? 2 : 1;
    long long M = (long long)g.B * g.Ti * upf * g.Hi * upf * g.Wi * upf;
    if (M == 0) return 0;
    const int ntaps_ = g.KT * g.KH * g.KW;
    const int act_mode = accumulate ? ACT_ACCUM : ACT_NONE;
    if (accumulate && (g.stride == 2 || g.Ci <= 4)) { rdg_set_error("simt_conv_bwd_data: accumulate is only supported by the generic path"); return -1; }
    if (!g.up && g.Ci <= 4 && (g.Co & 3) == 0 && (size_t)ntaps_ * g.Ci * g.Co * 4 <= 48 * 1024) {
        conv_bwd_data_smallci_kernel<<<ceil_div(M * g.Ci, 256), 256, (size_t)ntaps_ * g.Ci * g.Co * 4, st>>>(dy, w, dx, g, g.Ci);
        RDG_LAUNCH_CHECK();
        return 0;
    }
    if (!g.up && g.stride == 2) {
        // parity classes; the slice count follows the largest class (8 of 27 taps) at the nominal batch, see simt_conv_fwd
        ClsInfo ci{};
        const int Bn = g.B < 32 ? 32 : g.B;
        int tiles = 0, tiles_nom = 0;
        for (int cls = 0; cls < 8; ++cls) {
            const int qt = cls >> 2, qh = (cls >> 1) & 1, qw = cls & 1;
            const int ot = (qt - g.pt) & 1, oh = (qh - g.ph) & 1, ow = (qw - g.pw) & 1;
            const long long per = (long long)((g.Ti - ot + 1) / 2) * ((g.Hi - oh + 1) / 2) * ((g.Wi - ow + 1) / 2);
            ci.tile_begin[cls] = tiles;
            tiles += ceil_div(per * g.B, BM);
            tiles_nom += ceil_div(per * Bn, BM);
        }
        ci.tile_begin[8] = tiles;
        dim3 grid(tiles, ceil_div(g.Ci, BN));
        const int steps_max = ((g.KT + 1) / 2) * ((g.KH + 1) / 2) * ((g.KW + 1) / 2) * ceil_div(g.Co, BK);
        const int nslice = splitk_slices(tiles_nom * grid.y, steps_max);
        if (nslice > 1) {
            float* part = nullptr;
            RDG_CUDA(cudaMallocAsync(&part, (size_t)nslice * M * g.Ci * sizeof(float), st));
            grid.z = nslice;
            conv_bwd_data_s2_kernel<<<grid, NT, 0, st>>>(dy, w, dx, g, ci, part);
            RDG_LAUNCH_CHECK();
            splitk_epilogue_kernel<<<ceil_div(M * g.Ci, 256), 256, 0, st>>>(part, nslice, M * g.Ci, g.Ci, nullptr, dx, ACT_NONE, nullptr, 1.f, nullptr);
            RDG_LAUNCH_CHECK();
            RDG_CUDA(cudaFreeAsync(part, st));
            return 0;
        }
        conv_bwd_data_s2_kernel<<<grid, NT, 0, st>>>(dy, w, dx, g, ci, nullptr);
        RDG_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid(ceil_div(M, BM), ceil_div(g.Ci, BN));
    const int nslice = splitk_slices(ceil_div(M / g.B * (g.B < 32 ? 32 : g.B), BM) * grid.y, g.KT * g.KH * g.KW * ceil_div(g.Co, BK));
    if (nslice > 1) {
        float* part = nullptr;
        RDG_CUDA(cudaMallocAsync(&part, (size_t)nslice * M * g.Ci * sizeof(float), st));
        grid.z = nslice;
        conv_gemm_kernel<1><<<grid, NT, 0, st>>>(dy, w, nullptr, dx, g, ACT_NONE, nullptr, 1.f, nullptr, part);
        RDG_LAUNCH_CHECK();
        splitk_epilogue_kernel<<<ceil_div(M * g.Ci, 256), 256, 0, st>>>(part, nslice, M * g.Ci, g.Ci, nullptr, dx, act_mode, nullptr, 1.f, nullptr);
        RDG_LAUNCH_CHECK();
        RDG_CUDA(cudaFreeAsync(part, st));
        return 0;
    }
    conv_gemm_kernel<1><<<grid, NT, 0, st>>>(dy, w, nullptr, dx, g, act_mode, nullptr, 1.f, nullptr, nullptr);
    RDG_LAUNCH_CHECK();
    return 0;
}

// gradient w.r.t. input channel 0 only (dx keeps its Ci-channel layout; channels >= 1 are not written): the critic's first conv in
// the training steps, where only the sample channel's gradient is used (gradient-penalty norm :238-241, generator step)
//
// With a scratch buffer and Co = 64 the work is done as tap products: P[tap][pos] = <dy[pos][:], w[tap][0][:]> for every output
// position and tap (one small dense pass, dy read once, coalesced), then every input pixel sums the <= 8 products that reach it.
// The direct form below walks 8 taps x 64 channels of strided global loads per thread and took 55 us at batch 32, on the
// critical path of every step; this form takes two short launches.
__global__ void __launch_bounds__(256) ch0_tap_products_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                               float* __restrict__ P, long long M1, int Ci, int ntaps) {
    __shared__ float sa[64][65];
    __shared__ __align__(16) float sw[32][64];
    const long long pos0 = (long long)blockIdx.x * 64;
    for (int i = threadIdx.x; i < 1024; i += 256) {
        const int r = i >> 4, c4 = i & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pos0 + r < M1) v = __ldg(reinterpret_cast<const float4*>(dy + (pos0 + r) * 64) + c4);
        sa[r][c4 * 4 + 0] = v.x; sa[r][c4 * 4 + 1] = v.y; sa[r][c4 * 4 + 2] = v.z; sa[r][c4 * 4 + 3] = v.w;
    }
    for (int i = threadIdx.x; i < ntaps * 64; i += 256) sw[i >> 6][i & 63] = w[(size_t)(i >> 6) * Ci * 64 + (i & 63)];
    __syncthreads();
    const int pos = threadIdx.x & 63, grp = threadIdx.x >> 6;        // a warp shares its taps: weight reads are broadcasts
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int c4 = 0; c4 < 16; ++c4) {
        const float a0 = sa[pos][c4 * 4], a1 = sa[pos][c4 * 4 + 1], a2 = sa[pos][c4 * 4 + 2], a3 = sa[pos][c4 * 4 + 3];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int tap = grp + 4 * j;
            if (tap < ntaps) {
                const float4 b = *reinterpret_cast<const float4*>(&sw[tap][c4 * 4]);
                acc[j] = fmaf(a0, b.x, fmaf(a1, b.y, fmaf(a2, b.z, fmaf(a3, b.w, acc[j]))));
            }
        }
    }
    if (pos0 + pos < M1) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (grp + 4 * j < ntaps) P[(long long)(grp + 4 * j) * M1 + pos0 + pos] = acc[j];
    }
}

__global__ void __launch_bounds__(256) ch0_gather_kernel(const float* __restrict__ P, float* __restrict__ dx, ConvGeom g, long long M1) {
    // 32-bit index arithmetic (the host checks the sizes); per axis only the taps k = (p + pad) mod stride, + stride, ... reach p
    const int M = g.B * g.Ti * g.Hi * g.Wi;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M) return;
    int r = idx;
    const int w = r % g.Wi; r /= g.Wi;
    const int h = r % g.Hi; r /= g.Hi;
    const int t = r % g.Ti; const int b = r / g.Ti;
    const int s = g.stride;
    float acc = 0.f;
    for (int kt = (t + g.pt) % s; kt < g.KT; kt += s) {
        const int nt = (t + g.pt - kt) / s;
        if (t + g.pt - kt < 0 || nt >= g.To) continue;
        for (int kh = (h + g.ph) % s; kh < g.KH; kh += s) {
            const int nh = (h + g.ph - kh) / s;
            if (h + g.ph - kh < 0 || nh >= g.Ho) continue;
            const int rowbase = ((b * g.To + nt) * g.Ho + nh) * g.Wo;
            for (int kw = (w + g.pw) % s; kw < g.KW; kw += s) {
                const int nw = (w + g.pw - kw) / s;
                if (w + g.pw - kw < 0 || nw >= g.Wo) continue;
                acc += __ldg(P + (long long)((kt * g.KH + kh) * g.KW + kw) * M1 + rowbase + nw);
            }
        }
    }
    dx[(long long)idx * g.Ci] = acc;
}

int simt_conv_bwd_data_ch0(const float* dy, const float* w, float* dx, const ConvGeom& g, cudaStream_t st, float* scratch) {
    const long long M = (long long)g.B * g.Ti * g.Hi * g.Wi;
    const int ntaps_ = g.KT * g.KH * g.KW;
    if (M == 0) return 0;
    if (scratch && !g.up && g.Co == 64 && ntaps_ <= 32 && M * g.Ci < (1ll << 31)) {
        const long long M1 = (long long)g.B * g.To * g.Ho * g.Wo;
        ch0_tap_products_kernel<<<ceil_div(M1, 64), 256, 0, st>>>(dy, w, scratch, M1, g.Ci, ntaps_);
        RDG_LAUNCH_CHECK();
        ch0_gather_kernel<<<ceil_div(M, 256), 256, 0, st>>>(scratch, dx, g, M1);
        RDG_LAUNCH_CHECK();
        return 0;
    }
    if (g.up || g.Ci > 4 || (g.Co & 3) || (size_t)ntaps_ * g.Ci * g.Co * 4 > 48 * 1024) return simt_conv_bwd_data(dy, w, dx, g, st, 0);
    conv_bwd_data_smallci_kernel<<<ceil_div(M, 256), 256, (size_t)ntaps_ * g.Ci * g.Co * 4, st>>>(dy, w, dx, g, 1);
    RDG_LAUNCH_CHECK();
    return 0;
}

int simt_conv_bwd_filter(const float* x, const float* dy, float* dw, float* db, const ConvGeom& g,
                         cudaStream_t st) {
    long long M = (long long)g.B * g.To * g.Ho * g.Wo;
    if (M == 0) return 0;
    int ntaps = g.KT * g.KH * g.KW;
    if (!g.up && (g.Ci <= 4 || g.Co <= 4) && g.Ci * g.Co <= 256) {
        const int nsub = 512 / (g.Ci * g.Co);
        int nslice = ceil_div(M, (long long)64 * nsub);             // ~64 rows per thread ...
        if (nslice > 128) nslice = 128;
        if (nslice < ceil_div(M, 2048)) nslice = ceil_div(M, 2048); // ... and at most 2048 rows per block (offset table)
        if ((long long)g.B * g.Ti * g.Hi * g.Wi * g.Ci >= (1ll << 31)) { rdg_set_error("conv_bwd_filter_small: tensor too large"); return -1; }
        conv_bwd_filter_small_kernel<<<dim3(ntaps, nslice), 512, 0, st>>>(x, dy, dw, g, nslice);
        RDG_LAUNCH_CHECK();
        if (db) return simt_colsum(dy, db, M, g.Co, st);
        return 0;
    }
    int tiles = ceil_div(g.Ci, BM) * ceil_div(g.Co, BN) * ntaps;
    int ksplit = 1;
    while (tiles * ksplit < 592 && (M / (ksplit * 2)) >= 256) ksplit *= 2;
    dim3 grid(ceil_div(g.Ci, BM), ceil_div(g.Co, BN), ntaps * ksplit);
    conv_bwd_filter_kernel<<<grid, NT, 0, st>>>(x, dy, dw, g, ksplit);
    RDG_LAUNCH_CHECK();
    if (db) return simt_colsum(dy, db, M, g.Co, st);
    return 0;
}

int simt_colsum(const float* x, float* out, long long rows, int C, cudaStream_t st) {
    if (rows <= 0) return 0;
    long long rpb = 256;
    colsum_kernel<<<dim3(ceil_div(rows, rpb), ceil_div(C, 32)), 256, 0, st>>>(x, out, rows, C, rpb);
    RDG_LAUNCH_CHECK();
    return 0;
}
