// Elementwise / reduction kernels of the RainDisaggGAN hot path (FP32).
//   PixelNormalization   gan_train_cwgangp_pixelnorm.py:249-270
//   LeakyReLU(0.2)       :327,333,338,343 / :288-300
//   Softmax(axis=1)      :347  (+ check_numerics :349-350, + mm rescale raindisagg_gan_pretrained.py:62-64)
//   cond tiling/concat   :275-282 ; latent/cond concat :319-323
//   Adam                 :385 (Keras OptimizerV2 form)
#include "rdg_common.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void assemble_gen_input_kernel(const float* __restrict__ latent, const float* __restrict__ cond,
                                          int spc, int b_off, float* __restrict__ x0, int B, int ncf) {
    const int K = RDG_LATENT + ncf;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * K) return;
    int b = (int)(i / K), k = (int)(i % K);
    x0[i] = k < RDG_LATENT ? latent[(long long)b * RDG_LATENT + k]
                           : cond[(long long)((b_off + b) / spc) * ncf + (k - RDG_LATENT)];
}

// one warp per row of C channels
__global__ void pixelnorm_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows, int C,
                                 int lrelu) {
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* xr = x + row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) { float v = xr[c]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    const float l2 = sqrtf(s / (float)C + 1.0e-8f);
    for (int c = lane; c < C; c += 32) {
        float v = xr[c] / l2;
        if (lrelu) v = v > 0.f ? v : 0.2f * v;
        y[row * C + c] = v;
    }
}

// backward of LeakyReLU(PixelNorm(x)) given the conv output x (pre-norm) -- SURVEY A4
__global__ void pixelnorm_lrelu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                           float* __restrict__ dx, long long rows, int C) {
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* xr = x + row * C;
    const float* gr = dy + row * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) { float v = xr[c]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    const float l2 = sqrtf(s / (float)C + 1.0e-8f);
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) {
        float yn = xr[c] / l2;
        float g = gr[c] * (yn > 0.f ? 1.f : 0.2f);
        dot = fmaf(g, yn, dot);
    }
    dot = warp_sum(dot) / (float)C;
    for (int c = lane; c < C; c += 32) {
        float yn = xr[c] / l2;
        float g = gr[c] * (yn > 0.f ? 1.f : 0.2f);
        dx[row * C + c] = (g - yn * dot) / l2;
    }
}

// The same two functions for C = 64 / 128 / 256 (every PixelNorm of the generator) with the row held in registers: 16-byte
// loads, C/4 (at most 32) lanes per row, so a 64-channel row takes half a warp and nothing is read twice.  The scalar kernels
// above ran at a third of the HBM/L2 rate on the 50 MB activations of a batch-32 training step.
template <int L>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int C>
__global__ void __launch_bounds__(256) pixelnorm_vec_kernel(const float* __restrict__ x, float* __restrict__ y, long long rows, int lrelu) {
    constexpr int L = C / 4 < 32 ? C / 4 : 32, V = C / (4 * L), RPW = 32 / L;
    const int lane = threadIdx.x & 31, l = lane % L;
    const long long row = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + lane / L;
    const bool ok = row < rows;
    const float4* xr = reinterpret_cast<const float4*>(x) + (ok ? row : 0) * (C / 4);
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        v[j] = xr[j * L + l];
        s = fmaf(v[j].x, v[j].x, fmaf(v[j].y, v[j].y, fmaf(v[j].z, v[j].z, fmaf(v[j].w, v[j].w, s))));
    }
    s = group_sum<L>(s);
    if (!ok) return;
    const float l2 = sqrtf(s / (float)C + 1.0e-8f);
    float4* yr = reinterpret_cast<float4*>(y) + row * (C / 4);
#pragma unroll
    for (int j = 0; j < V; ++j) {
        float4 o = make_float4(v[j].x / l2, v[j].y / l2, v[j].z / l2, v[j].w / l2);
        if (lrelu) {
            o.x = o.x > 0.f ? o.x : 0.2f * o.x; o.y = o.y > 0.f ? o.y : 0.2f * o.y;
            o.z = o.z > 0.f ? o.z : 0.2f * o.z; o.w = o.w > 0.f ? o.w : 0.2f * o.w;
        }
        yr[j * L + l] = o;
    }
}

template <int C>
__global__ void __launch_bounds__(256) pixelnorm_lrelu_bwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                      float* __restrict__ dx, long long rows) {
    constexpr int L = C / 4 < 32 ? C / 4 : 32, V = C / (4 * L), RPW = 32 / L;
    const int lane = threadIdx.x & 31, l = lane % L;
    const long long row = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + lane / L;
    const bool ok = row < rows;
    const float4* xr = reinterpret_cast<const float4*>(x) + (ok ? row : 0) * (C / 4);
    const float4* gr = reinterpret_cast<const float4*>(dy) + (ok ? row : 0) * (C / 4);
    float4 v[V], g[V];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        v[j] = xr[j * L + l];
        g[j] = gr[j * L + l];
        s = fmaf(v[j].x, v[j].x, fmaf(v[j].y, v[j].y, fmaf(v[j].z, v[j].z, fmaf(v[j].w, v[j].w, s))));
    }
    s = group_sum<L>(s);
    const float l2 = sqrtf(s / (float)C + 1.0e-8f);
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j) {
        v[j].x /= l2; v[j].y /= l2; v[j].z /= l2; v[j].w /= l2;              // yn
        g[j].x *= v[j].x > 0.f ? 1.f : 0.2f; g[j].y *= v[j].y > 0.f ? 1.f : 0.2f;
        g[j].z *= v[j].z > 0.f ? 1.f : 0.2f; g[j].w *= v[j].w > 0.f ? 1.f : 0.2f;
        dot = fmaf(g[j].x, v[j].x, fmaf(g[j].y, v[j].y, fmaf(g[j].z, v[j].z, fmaf(g[j].w, v[j].w, dot))));
    }
    dot = group_sum<L>(dot) / (float)C;
    if (!ok) return;
    float4* dr = reinterpret_cast<float4*>(dx) + row * (C / 4);
#pragma unroll
    for (int j = 0; j < V; ++j)
        dr[j * L + l] = make_float4((g[j].x - v[j].x * dot) / l2, (g[j].y - v[j].y * dot) / l2, (g[j].z - v[j].z * dot) / l2,
                                    (g[j].w - v[j].w * dot) / l2);
}

// logits [B,24,P] -> fractions (or mm) [B,24,P]; thread per (b,p)
__global__ void softmax_hours_kernel(const float* __restrict__ logits, float* __restrict__ out, long long B, int P,
                                     const float* __restrict__ cond, int spc, int ncond, float scale, int out_mm,
                                     int* __restrict__ nonfinite) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * P) return;
    long long b = i / P; int p = (int)(i % P);
    const float* l = logits + b * RDG_NHOURS * P + p;
    float v[RDG_NHOURS], mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < RDG_NHOURS; ++t) { v[t] = l[(long long)t * P]; mx = fmaxf(mx, v[t]); }
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < RDG_NHOURS; ++t) { v[t] = expf(v[t] - mx); s += v[t]; }
    float mul = 1.f;
    if (out_mm) mul = cond[((b / spc) * P + p) * ncond] * scale;
    bool bad = false;
#pragma unroll
    for (int t = 0; t < RDG_NHOURS; ++t) {
        float f = v[t] / s;
        bad |= !isfinite(f);
        out[b * RDG_NHOURS * P + (long long)t * P + p] = f * mul;
    }
    if (bad && nonfinite) atomicOr(nonfinite, 1);
}

__global__ void softmax_hours_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                         float* __restrict__ dl, long long B, int P) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * P) return;
    long long b = i / P; int p = (int)(i % P);
    long long base = b * RDG_NHOURS * P + p;
    float dot = 0.f;
    for (int t = 0; t < RDG_NHOURS; ++t) dot = fmaf(y[base + (long long)t * P], dy[base + (long long)t * P], dot);
    for (int t = 0; t < RDG_NHOURS; ++t) {
        long long k = base + (long long)t * P;
        dl[k] = y[k] * (dy[k] - dot);
    }
}

__global__ void lrelu_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ dy, float* __restrict__ dx,
                                 long long n, const float* __restrict__ mask, float mask_scale) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float g = dy[i] * (pre[i] > 0.f ? 1.f : 0.2f);
    if (mask) g *= mask[i] * mask_scale;
    dx[i] = g;
}

// gradient of nearest x2 upsample: 2x2x2 sum-pool
__global__ void upsample_pool_kernel(const float* __restrict__ du, float* __restrict__ dl, int B, int T, int H, int W,
                                     int C) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long n = (long long)B * T * H * W * C;
    if (i >= n) return;
    int c = (int)(i % C); long long r = i / C;
    int w = (int)(r % W); r /= W;
    int h = (int)(r % H); r /= H;
    int t = (int)(r % T); int b = (int)(r / T);
    float s = 0.f;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        int tt = 2 * t + (a >> 2), hh = 2 * h + ((a >> 1) & 1), ww = 2 * w + (a & 1);
        s += du[((((long long)b * 2 * T + tt) * 2 * H + hh) * 2 * W + ww) * C + c];
    }
    dl[i] = s;
}

__global__ void critic_input_kernel(const float* __restrict__ sample, const float* __restrict__ cond,
                                    float* __restrict__ x, int B, int nd, int ncond) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int C = 1 + ncond;
    long long n = (long long)B * RDG_NHOURS * nd * nd * C;
    if (i >= n) return;
    int c = (int)(i % C); long long r = i / C;
    int p = (int)(r % (nd * nd)); r /= (nd * nd);
    long long b = r / RDG_NHOURS;
    x[i] = c == 0 ? sample[i / C] : cond[(b * nd * nd + p) * ncond + (c - 1)];
}

// The three critic inputs of a critic step in one pass (gan_train_cwgangp_pixelnorm.py:221-224, :275-282, :372-379):
// x3 = [fake | real | alpha*real + (1-alpha)*fake], each [B,24,nd,nd,1+ncond] with the condition tiled over the hours
__global__ void critic_inputs3_kernel(const float* __restrict__ fake, const float* __restrict__ real, const float* __restrict__ alpha,
                                      const float* __restrict__ cond, float* __restrict__ x3, int B, int nd, int ncond) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (b, t, p)
    const int P = nd * nd;
    const long long n = (long long)B * RDG_NHOURS * P;
    if (i >= n) return;
    const int p = (int)(i % P);
    const long long b = i / ((long long)RDG_NHOURS * P);
    const int C = 1 + ncond;
    const float f = fake[i], r = real[i], a = alpha[b];
    const float v[3] = {f, r, a * r + (1.f - a) * f};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float* d = x3 + ((long long)k * n + i) * C;
        d[0] = v[k];
        for (int c = 0; c < ncond; ++c) d[1 + c] = cond[(b * P + p) * ncond + c];
    }
}

// Philox4x32-10 counter-based generator + Box-Muller; element i uses counter (offset+i)/4.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}

__global__ void fill_normal_kernel(float* __restrict__ dst, long long n, uint64_t seed, uint64_t offset) {
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // quad index
    if (q * 4 >= n) return;
    uint64_t ctr = offset / 4 + (uint64_t)q;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c, k);
    float u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = ((float)c[j] + 0.5f) * 2.3283064365386963e-10f;   // (0,1)
    float r0 = sqrtf(-2.f * logf(u[0])), r1 = sqrtf(-2.f * logf(u[2]));
    float s0, c0, s1, c1;
    sincospif(2.f * u[1], &s0, &c0);
    sincospif(2.f * u[3], &s1, &c1);
    float z[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
#pragma unroll
    for (int j = 0; j < 4; ++j) if (q * 4 + j < n) dst[q * 4 + j] = z[j];
}

// ---- device-resident training state: lets a captured CUDA graph of a training step be replayed (the Philox counter and the
// Adam step live in device memory and are advanced by kernels inside the graph instead of being baked into kernel arguments)
__global__ void train_tick_kernel(RdgTrainState* s, int which) { s->rng_ctr[which] += 1ull; }
__global__ void adam_prepare_kernel(RdgTrainState* s, float lr, float b1, float b2) {
    const long long t = ++s->adam_t;      // shared step counter of the one optimizer object (gan_train_cwgangp_pixelnorm.py:385)
    s->lr_t = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
}
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                long long n4, const RdgTrainState* __restrict__ s, float b1, float b2, float eps, float gs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float lr_t = s->lr_t;
    float4 gi = reinterpret_cast<const float4*>(g)[i], mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i];
    float4 pi = reinterpret_cast<float4*>(p)[i];
    float* G = &gi.x; float* M = &mi.x; float* V = &vi.x; float* P = &pi.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float gj = G[j] * gs;
        M[j] = b1 * M[j] + (1.f - b1) * gj;
        V[j] = b2 * V[j] + (1.f - b2) * gj * gj;
        P[j] = P[j] - lr_t * M[j] / (sqrtf(V[j]) + eps);
    }
    reinterpret_cast<float4*>(m)[i] = mi; reinterpret_cast<float4*>(v)[i] = vi; reinterpret_cast<float4*>(p)[i] = pi;
}
// kind 0: N(0,1) (Box-Muller), 1: U[0,1), 2: Bernoulli(keep) as 0/1 floats.  Philox counter = (quad index, stream id, step counter)
__global__ void fill_random_dev_kernel(float* __restrict__ dst, long long n, uint64_t seed, const RdgTrainState* __restrict__ s,
                                       int which, uint32_t stream_id, int kind, float keep) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q * 4 >= n) return;
    uint32_t c[4] = {(uint32_t)q, (uint32_t)((uint64_t)q >> 32), stream_id + 16u * (uint32_t)which, (uint32_t)s->rng_ctr[which]};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c, k);
    float z[4];
    if (kind == 0) {
        float u[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = ((float)c[j] + 0.5f) * 2.3283064365386963e-10f;
        const float r0 = sqrtf(-2.f * logf(u[0])), r1 = sqrtf(-2.f * logf(u[2]));
        float s0, c0, s1, c1;
        sincospif(2.f * u[1], &s0, &c0);
        sincospif(2.f * u[3], &s1, &c1);
        z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float u = (float)(c[j] >> 8) * 5.9604644775390625e-08f;     // [0,1) with 24 bits
            z[j] = kind == 1 ? u : (u < keep ? 1.f : 0.f);
        }
    }
    if (q * 4 + 3 < n) *reinterpret_cast<float4*>(dst + q * 4) = make_float4(z[0], z[1], z[2], z[3]);
    else for (int j = 0; q * 4 + j < n; ++j) dst[q * 4 + j] = z[j];
}

__device__ __forceinline__ void philox_draw(uint32_t (&c)[4], uint64_t seed) {
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c, k);
}
__global__ void step_random_dev_kernel(float* __restrict__ buf, long long qa, long long n_normal, long long off_u, long long qb, long long n_u,
                                       long long off_m, long long n_m, float keep, uint64_t seed, RdgTrainState* s, int which) {
    // a bounded grid walks the quads (one same-address atomic per BLOCK below: 4700 one-quad blocks made that the bottleneck, 33 us)
    __shared__ uint32_t s_ctr;
    if (threadIdx.x == 0) s_ctr = (uint32_t)(*reinterpret_cast<volatile unsigned long long*>(&s->rng_ctr[which]) + 1ull);
    __syncthreads();
    const uint32_t ctr = s_ctr;
    for (long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < qa + qb + (n_m + 3) / 4; q0 += (long long)gridDim.x * blockDim.x) {
    long long q = q0;
    float* dst; long long n; uint32_t sid; int kind;
    if (q < qa) { dst = buf; n = n_normal; sid = 1; kind = 0; }
    else if (q < qa + qb) { q -= qa; dst = buf + off_u; n = n_u; sid = 2; kind = 1; }
    else { q -= qa + qb; dst = buf + off_m; n = n_m; sid = 3; kind = 2; }
    if (q * 4 < n) {
        uint32_t c[4] = {(uint32_t)q, (uint32_t)((uint64_t)q >> 32), sid + 16u * (uint32_t)which, ctr};
        philox_draw(c, seed);
        float z[4];
        if (kind == 0) {
            float u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) u[j] = ((float)c[j] + 0.5f) * 2.3283064365386963e-10f;
            const float r0 = sqrtf(-2.f * logf(u[0])), r1 = sqrtf(-2.f * logf(u[2]));
            float s0, c0, s1, c1;
            sincospif(2.f * u[1], &s0, &c0);
            sincospif(2.f * u[3], &s1, &c1);
            z[0] = r0 * c0; z[1] = r0 * s0; z[2] = r1 * c1; z[3] = r1 * s1;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float u = (float)(c[j] >> 8) * 5.9604644775390625e-08f;
                z[j] = kind == 1 ? u : (u < keep ? 1.f : 0.f);
            }
        }
        if (q * 4 + 3 < n) *reinterpret_cast<float4*>(dst + q * 4) = make_float4(z[0], z[1], z[2], z[3]);
        else for (int j = 0; q * 4 + j < n; ++j) dst[q * 4 + j] = z[j];
    }
    }
    // the last block to finish advances the step counter (every block read the old value before it arrived here)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&s->done[which], 1u) == gridDim.x - 1) {
            s->done[which] = 0u;
            s->rng_ctr[which] += 1ull;
            __threadfence();
        }
    }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr_t, float b1, float b2, float eps,
                            float gs) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = g[i] * gs;
    float mi = b1 * m[i] + (1.f - b1) * gi;
    float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
}

__global__ void interp_kernel(const float* __restrict__ xr, const float* __restrict__ xf, const float* __restrict__ alpha,
                              float* __restrict__ xh, long long n, long long per) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = alpha[i / per];
    xh[i] = a * xr[i] + (1.f - a) * xf[i];          // RandomWeightedAverage, gan_train_cwgangp_pixelnorm.py:221-224
}
// Flatten + Dense(1) (gan_train_cwgangp_pixelnorm.py:303-304): one warp per sample, score[b] = x[b,:] . w + bias
__global__ void dense_score_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                   float* __restrict__ score, int B, int K) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)b * K);
    const float4* wr = reinterpret_cast<const float4*>(w);
    float s = 0.f;
    for (int i = lane; i < K / 4; i += 32) {
        const float4 a = xr[i], c = __ldg(wr + i);
        s = fmaf(a.x, c.x, fmaf(a.y, c.y, fmaf(a.z, c.z, fmaf(a.w, c.w, s))));
    }
    s = warp_sum(s);
    if (lane == 0) score[b] = s + bias[0];
}
// Tail of the critic forward / head of its backward in one launch (tensor-core training mode): the per-sample cotangents of the
// scores dscore[b] = cot[b / B] (segments of B samples: e.g. [+1/B | -1/B | 1] for fake | real | interpolated, :452-454, :238-241),
// the Wasserstein loss terms loss_k = sign_k * mean(score of segment k) (k < nloss), and the cotangent of the last conv's
// pre-activation da4[b][k] = dscore[b] * W5[k] * LeakyReLU'(a4[b][k]) * mask (the Dense(1) backward + LeakyReLU/dropout backward).
__global__ void critic_tail_kernel(const float* __restrict__ score, const float* __restrict__ w5, const float* __restrict__ a4,
                                   const float* __restrict__ mask, float mask_scale, int B, int nseg, int K, float c0, float c1, float c2,
                                   float s0, float s1, int nloss, float* __restrict__ loss, float* __restrict__ dscore,
                                   float* __restrict__ da4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // float4 index into [nseg*B][K]
    const long long n4 = (long long)nseg * B * K / 4;
    if (i < n4) {
        const long long e = i * 4;
        const int b = (int)(e / K), k = (int)(e - (long long)b * K);
        const int seg = b / B;
        const float ct = seg == 0 ? c0 : (seg == 1 ? c1 : c2);
        const float4 w = __ldg(reinterpret_cast<const float4*>(w5 + k));
        const float4 a = reinterpret_cast<const float4*>(a4)[i];
        float4 r = make_float4(ct * w.x * (a.x > 0.f ? 1.f : 0.2f), ct * w.y * (a.y > 0.f ? 1.f : 0.2f),
                               ct * w.z * (a.z > 0.f ? 1.f : 0.2f), ct * w.w * (a.w > 0.f ? 1.f : 0.2f));
        if (mask) {
            const float4 m = reinterpret_cast<const float4*>(mask)[i];
            r.x *= m.x * mask_scale; r.y *= m.y * mask_scale; r.z *= m.z * mask_scale; r.w *= m.w * mask_scale;
        }
        reinterpret_cast<float4*>(da4)[i] = r;
        if (k == 0 && dscore) dscore[b] = ct;
    }
    if (blockIdx.x == 0) {                       // the loss terms: one warp per term, fixed summation order
        const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (wp < nloss) {
            float s = 0.f;
            for (int j = lane; j < B; j += 32) s += score[wp * B + j];
            s = warp_sum(s);
            if (lane == 0) loss[wp] = (wp == 0 ? s0 : s1) * s / (float)B;
        }
    }
}

// ---- output conv Conv3D(64->1,'same') through per-tap products (training, tensor-core mode): P[pos][tap] = y[pos] . w4[tap]
// logits[b,t,h,w] = b4 + sum_tap P[(t+kt-1, h+kh-1, w+kw-1)][tap]       (gan_train_cwgangp_pixelnorm.py:345)
__global__ void tap_gather_logits_kernel(const float* __restrict__ P, const float* __restrict__ b4, float* __restrict__ logits,
                                         int B, int nd) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)B * RDG_NHOURS * nd * nd;
    if (i >= n) return;
    const int w = (int)(i % nd), h = (int)((i / nd) % nd), t = (int)((i / ((long long)nd * nd)) % RDG_NHOURS);
    float s = b4[0];
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
        const int tt = t + kt - 1;
        if ((unsigned)tt >= (unsigned)RDG_NHOURS) continue;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int hh = h + kh - 1;
            if ((unsigned)hh >= (unsigned)nd) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int ww = w + kw - 1;
                if ((unsigned)ww >= (unsigned)nd) continue;
                const long long q = i + ((long long)(kt - 1) * nd + (kh - 1)) * nd + (kw - 1);
                s += P[q * 32 + (kt * 3 + kh) * 3 + kw];
            }
        }
    }
    logits[i] = s;
}
// its transpose: Gd[pos][tap] = dlogits[pos - (k - 1)] (0 outside the grid, taps 27..31 zero) so that
// dy = Gd . w4 and dw4 = Gd^T . y are plain GEMMs
__global__ void tap_scatter_dlogits_kernel(const float* __restrict__ dl, float* __restrict__ Gd, int B, int nd) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // (position, tap quad)
    const long long n = (long long)B * RDG_NHOURS * nd * nd;
    if (i >= n * 8) return;
    const long long pos = i >> 3;
    const int tq = (int)(i & 7);
    const int w = (int)(pos % nd), h = (int)((pos / nd) % nd), t = (int)((pos / ((long long)nd * nd)) % RDG_NHOURS);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int tap = tq * 4 + j;
        v[j] = 0.f;
        if (tap < 27) {
            const int kt = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
            const int tt = t - (kt - 1), hh = h - (kh - 1), ww = w - (kw - 1);
            if ((unsigned)tt < (unsigned)RDG_NHOURS && (unsigned)hh < (unsigned)nd && (unsigned)ww < (unsigned)nd)
                v[j] = dl[pos - (((long long)(kt - 1) * nd + (kh - 1)) * nd + (kw - 1))];
        }
    }
    *reinterpret_cast<float4*>(Gd + i * 4) = make_float4(v[0], v[1], v[2], v[3]);
}
// w4 (27,64,1) -> w4p [32][64] (rows 27..31 zero) and, behind it, w4pT [64][32]
__global__ void pad_w4_kernel(const float* __restrict__ w4, float* __restrict__ w4p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 32 * 64) return;
    const int tap = i >> 6, ch = i & 63;
    const float v = tap < 27 ? w4[tap * 64 + ch] : 0.f;
    w4p[i] = v;
    w4p[2048 + ch * 32 + tap] = v;
}
// three constant segments of n each (the cotangents of the merged [fake | real | interpolated] critic backward)
__global__ void fill3_kernel(float* __restrict__ d, int n, float a, float b, float c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3 * n) d[i] = i < n ? a : (i < 2 * n ? b : c);
}
__global__ void fill_kernel(float* __restrict__ d, long long n, float v) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = v;
}
__device__ __forceinline__ float block_sum(float v) {
    __shared__ float sh[32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;   // valid in warp 0
}
__global__ void mean_scaled_kernel(const float* __restrict__ x, long long n, float scale, float* __restrict__ out) {
    float s = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
    s = block_sum(s);
    if (threadIdx.x == 0) out[0] = scale * s / (float)n;
}
// one block per sample: || g0[b, :, channel 0] ||_2   (GradientPenalty, gan_train_cwgangp_pixelnorm.py:240-241)
__global__ void gp_norm_kernel(const float* __restrict__ g0, int C, long long per, float* __restrict__ norm) {
    const float* g = g0 + (long long)blockIdx.x * per * C;
    float s = 0.f;
    for (long long i = threadIdx.x; i < per; i += blockDim.x) { float v = g[i * C]; s = fmaf(v, v, s); }
    s = block_sum(s);
    if (threadIdx.x == 0) norm[blockIdx.x] = sqrtf(s);      // no epsilon, like the reference
}
__global__ void gp_cotangent_kernel(const float* __restrict__ g0, const float* __restrict__ norm, float coef,
                                    float* __restrict__ u0, int C, long long n, long long per, int B, float* __restrict__ loss_out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (loss_out && blockIdx.x == 0 && threadIdx.x < 32) {       // 'mse' against zeros of the penalty output: mean((norm - 1)^2)
        float s = 0.f;
        for (int b = threadIdx.x; b < B; b += 32) { const float d = norm[b] - 1.f; s = fmaf(d, d, s); }
        s = warp_sum(s);
        if (threadIdx.x == 0) loss_out[0] = s / (float)B;
    }
    if (i >= n) return;
    const int c = (int)(i % C);
    const long long b = i / (per * C);
    const float nb = norm[b];
    u0[i] = c == 0 ? coef * (nb - 1.f) / nb * g0[i] : 0.f;
}
// norm + cotangent of one sample per block (tensor-core training mode: one launch on the critical chain instead of two); the
// penalty loss itself is formed from norm[] by combine_losses_norm_kernel
__global__ void __launch_bounds__(1024) gp_fused_kernel(const float* __restrict__ g0, float coef, float* __restrict__ u0, int C, long long per,
                                                        float* __restrict__ norm) {
    __shared__ float s_norm;
    const float* g = g0 + (long long)blockIdx.x * per * C;
    float* u = u0 + (long long)blockIdx.x * per * C;
    float s = 0.f;
    for (long long i = threadIdx.x; i < per; i += blockDim.x) { const float v = g[i * C]; s = fmaf(v, v, s); }
    s = block_sum(s);
    if (threadIdx.x == 0) { s_norm = sqrtf(s); norm[blockIdx.x] = s_norm; }      // no epsilon, like the reference
    __syncthreads();
    const float nb = s_norm, f = coef * (nb - 1.f) / nb;
    for (long long i = threadIdx.x; i < per * C; i += blockDim.x) u[i] = (i % C) == 0 ? f * g[i] : 0.f;
}
__global__ void combine_losses_norm_kernel(const float* lv, const float* lf, const float* __restrict__ norm, int B, float w, float* out4) {
    float s = 0.f;
    for (int b = threadIdx.x; b < B; b += 32) { const float d = norm[b] - 1.f; s = fmaf(d, d, s); }
    s = warp_sum(s);
    if (threadIdx.x == 0) {
        const float lgp = s / (float)B;            // 'mse' against zeros of the penalty output: mean((norm - 1)^2)
        out4[1] = lv[0]; out4[2] = lf[0]; out4[3] = lgp;
        out4[0] = lv[0] + lf[0] + w * lgp;         // loss_weights [1, 1, 10], gan_train_cwgangp_pixelnorm.py:388-392
    }
}
// loss_k = sign_k * mean(score of segment k): one warp per term, fixed summation order
__global__ void score_losses_kernel(const float* __restrict__ score, int B, int nloss, float s0, float s1, float* __restrict__ loss) {
    const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (wp >= nloss) return;
    float s = 0.f;
    for (int j = lane; j < B; j += 32) s += score[wp * B + j];
    s = warp_sum(s);
    if (lane == 0) loss[wp] = (wp == 0 ? s0 : s1) * s / (float)B;
}
__global__ void gp_loss_kernel(const float* __restrict__ norm, int B, float* __restrict__ out) {
    float s = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) { float d = norm[i] - 1.f; s = fmaf(d, d, s); }
    s = block_sum(s);
    if (threadIdx.x == 0) out[0] = s / (float)B;
}
__global__ void extract_channel0_kernel(const float* __restrict__ x, float* __restrict__ out, long long n, int C) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = x[i * C];
}
__global__ void combine_losses_kernel(const float* lv, const float* lf, const float* lgp, float w, float* out4) {
    out4[1] = lv[0]; out4[2] = lf[0]; out4[3] = lgp[0];
    out4[0] = lv[0] + lf[0] + w * lgp[0];     // loss_weights [1, 1, 10], gan_train_cwgangp_pixelnorm.py:388-392
}

}  // namespace

#define EW_GRID(n) ceil_div((n), 256), 256, 0, st

int ew_assemble_gen_input(const float* latent, const float* cond, int spc, int b_off, float* x0, int B, int ncf,
                          cudaStream_t st) {
    long long n = (long long)B * (RDG_LATENT + ncf);
    if (!n) return 0;
    assemble_gen_input_kernel<<<EW_GRID(n)>>>(latent, cond, spc, b_off, x0, B, ncf);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_pixelnorm(const float* x, float* y, long long rows, int C, int lrelu, cudaStream_t st) {
    if (!rows) return 0;
    const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    if (al && C == 64) pixelnorm_vec_kernel<64><<<ceil_div(rows, 16), 256, 0, st>>>(x, y, rows, lrelu);
    else if (al && C == 128) pixelnorm_vec_kernel<128><<<ceil_div(rows, 8), 256, 0, st>>>(x, y, rows, lrelu);
    else if (al && C == 256) pixelnorm_vec_kernel<256><<<ceil_div(rows, 8), 256, 0, st>>>(x, y, rows, lrelu);
    else pixelnorm_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(x, y, rows, C, lrelu);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_pixelnorm_lrelu_bwd(const float* x_pre, const float* dy, float* dx, long long rows, int C, cudaStream_t st) {
    if (!rows) return 0;
    const bool al = ((reinterpret_cast<uintptr_t>(x_pre) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
    if (al && C == 64) pixelnorm_lrelu_bwd_vec_kernel<64><<<ceil_div(rows, 16), 256, 0, st>>>(x_pre, dy, dx, rows);
    else if (al && C == 128) pixelnorm_lrelu_bwd_vec_kernel<128><<<ceil_div(rows, 8), 256, 0, st>>>(x_pre, dy, dx, rows);
    else if (al && C == 256) pixelnorm_lrelu_bwd_vec_kernel<256><<<ceil_div(rows, 8), 256, 0, st>>>(x_pre, dy, dx, rows);
    else pixelnorm_lrelu_bwd_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(x_pre, dy, dx, rows, C);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_softmax_hours(const float* logits, float* out, long long B, int P, const float* cond, int spc, int ncond,
                     float scale, int out_mm, int* nonfinite, cudaStream_t st) {
    if (!B) return 0;
    softmax_hours_kernel<<<EW_GRID(B * P)>>>(logits, out, B, P, cond, spc, ncond, scale, out_mm, nonfinite);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_softmax_hours_bwd(const float* y, const float* dy, float* dl, long long B, int P, cudaStream_t st) {
    if (!B) return 0;
    softmax_hours_bwd_kernel<<<EW_GRID(B * P)>>>(y, dy, dl, B, P);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_lrelu_bwd(const float* pre, const float* dy, float* dx, long long n, const float* mask, float mask_scale,
                 cudaStream_t st) {
    if (!n) return 0;
    lrelu_bwd_kernel<<<EW_GRID(n)>>>(pre, dy, dx, n, mask, mask_scale);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_upsample_pool(const float* d_up, float* d_lo, int B, int T, int H, int W, int C, cudaStream_t st) {
    long long n = (long long)B * T * H * W * C;
    if (!n) return 0;
    upsample_pool_kernel<<<EW_GRID(n)>>>(d_up, d_lo, B, T, H, W, C);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_critic_input(const float* sample, const float* cond, float* x, int B, int nd, int ncond, cudaStream_t st) {
    long long n = (long long)B * RDG_NHOURS * nd * nd * (1 + ncond);
    if (!n) return 0;
    critic_input_kernel<<<EW_GRID(n)>>>(sample, cond, x, B, nd, ncond);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_critic_inputs3(const float* fake, const float* real, const float* alpha, const float* cond, float* x3, int B, int nd, int ncond,
                      cudaStream_t st) {
    const long long n = (long long)B * RDG_NHOURS * nd * nd;
    if (!n) return 0;
    critic_inputs3_kernel<<<EW_GRID(n)>>>(fake, real, alpha, cond, x3, B, nd, ncond);
    RDG_LAUNCH_CHECK();
    return 0;
}

int ew_fill_normal(float* dst, long long n, uint64_t seed, uint64_t offset, cudaStream_t st) {
    if (!n) return 0;
    fill_normal_kernel<<<EW_GRID((n + 3) / 4)>>>(dst, n, seed, offset);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_adam(float* p, const float* g, float* m, float* v, long long n, float lr_t, float beta1, float beta2,
            float eps, float grad_scale, cudaStream_t st) {
    if (!n) return 0;
    adam_kernel<<<EW_GRID(n)>>>(p, g, m, v, n, lr_t, beta1, beta2, eps, grad_scale);
    RDG_LAUNCH_CHECK();
    return 0;
}

int ew_train_tick(RdgTrainState* s, int which, cudaStream_t st) {
    train_tick_kernel<<<1, 1, 0, st>>>(s, which);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_adam_dev(float* p, const float* g, float* m, float* v, long long n, RdgTrainState* s, float lr, float beta1, float beta2,
                float eps, float grad_scale, cudaStream_t st) {
    if (!n) return 0;
    if (n & 3) { rdg_set_error("ew_adam_dev: parameter count must be a multiple of 4 (flat buffers are padded)"); return -1; }
    adam_prepare_kernel<<<1, 1, 0, st>>>(s, lr, beta1, beta2);
    RDG_LAUNCH_CHECK();
    adam_dev_kernel<<<EW_GRID(n / 4)>>>(p, g, m, v, n / 4, s, beta1, beta2, eps, grad_scale);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_fill_random_dev(float* dst, long long n, uint64_t seed, const RdgTrainState* s, int which, uint32_t stream_id, int kind, float keep,
                       cudaStream_t st) {
    if (!n) return 0;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) != 0) { rdg_set_error("ew_fill_random_dev: destination must be 16-byte aligned"); return -1; }
    fill_random_dev_kernel<<<EW_GRID((n + 3) / 4)>>>(dst, n, seed, s, which, stream_id, kind, keep);
    RDG_LAUNCH_CHECK();
    return 0;
}

int ew_tap_gather_logits(const float* P, const float* b4, float* logits, int B, int nd, cudaStream_t st) {
    const long long n = (long long)B * RDG_NHOURS * nd * nd;
    if (!n) return 0;
    tap_gather_logits_kernel<<<EW_GRID(n)>>>(P, b4, logits, B, nd);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_tap_scatter_dlogits(const float* dl, float* Gd, int B, int nd, cudaStream_t st) {
    const long long n = (long long)B * RDG_NHOURS * nd * nd * 8;
    if (!n) return 0;
    tap_scatter_dlogits_kernel<<<EW_GRID(n)>>>(dl, Gd, B, nd);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_pad_w4(const float* w4, float* w4p, cudaStream_t st) {
    pad_w4_kernel<<<8, 256, 0, st>>>(w4, w4p);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_fill3(float* dst, int n, float a, float b, float c, cudaStream_t st) {
    if (!n) return 0;
    fill3_kernel<<<EW_GRID(3 * n)>>>(dst, n, a, b, c);
    RDG_LAUNCH_CHECK();
    return 0;
}

int ew_dense_score(const float* x, const float* w, const float* bias, float* score, int B, int K, cudaStream_t st) {
    if (!B) return 0;
    if (K & 3) { rdg_set_error("ew_dense_score: K must be a multiple of 4"); return -1; }
    dense_score_kernel<<<ceil_div(B, 8), 256, 0, st>>>(x, w, bias, score, B, K);
    RDG_LAUNCH_CHECK();
    return 0;
}

int ew_critic_tail(const float* score, const float* w5, const float* a4, const float* mask, float mask_scale, int B, int nseg, int K,
                   const float* cot3, const float* sign2, int nloss, float* loss, float* dscore, float* da4, cudaStream_t st) {
    if (K & 3) { rdg_set_error("ew_critic_tail: K must be a multiple of 4"); return -1; }
    const long long n4 = (long long)nseg * B * K / 4;
    critic_tail_kernel<<<EW_GRID(n4)>>>(score, w5, a4, mask, mask_scale, B, nseg, K, cot3[0], cot3[1], cot3[2], sign2[0], sign2[1], nloss,
                                          loss, dscore, da4);
    RDG_LAUNCH_CHECK();
    return 0;
}

int ew_step_random_dev(float* buf, long long n_normal, long long off_u, long long n_u, long long off_m, long long n_m, float keep, uint64_t seed,
                       RdgTrainState* s, int which, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(buf) & 15) || (off_u & 3) || (off_m & 3)) { rdg_set_error("ew_step_random_dev: regions must be 16-byte aligned"); return -1; }
    const long long qa = (n_normal + 3) / 4, qb = (n_u + 3) / 4, qc = (n_m + 3) / 4;
    const long long nblk = (qa + qb + qc + 255) / 256;
    step_random_dev_kernel<<<(unsigned)(nblk < 1184 ? nblk : 1184), 256, 0, st>>>(buf, qa, n_normal, off_u, qb, n_u, off_m, n_m, keep, seed, s, which);
    RDG_LAUNCH_CHECK();
    return 0;
}

int ew_interp(const float* xr, const float* xf, const float* alpha, float* xhat, int B, long long per, cudaStream_t st) {
    long long n = (long long)B * per;
    if (!n) return 0;
    interp_kernel<<<EW_GRID(n)>>>(xr, xf, alpha, xhat, n, per);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_fill(float* dst, long long n, float v, cudaStream_t st) {
    if (!n) return 0;
    fill_kernel<<<EW_GRID(n)>>>(dst, n, v);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_mean_scaled(const float* x, long long n, float scale, float* out, cudaStream_t st) {
    mean_scaled_kernel<<<1, 256, 0, st>>>(x, n, scale, out);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_gp_norm(const float* g0, int C, int B, long long per, float* norm, cudaStream_t st) {
    if (!B) return 0;
    gp_norm_kernel<<<B, 256, 0, st>>>(g0, C, per, norm);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_gp_cotangent(const float* g0, const float* norm, float coef, float* u0, int C, int B, long long per, cudaStream_t st, float* loss_out) {
    long long n = (long long)B * per * C;
    if (!n) return 0;
    gp_cotangent_kernel<<<EW_GRID(n)>>>(g0, norm, coef, u0, C, n, per, B, loss_out);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_gp_fused(const float* g0, float coef, float* u0, int C, int B, long long per, float* norm, cudaStream_t st) {
    if (!B) return 0;
    gp_fused_kernel<<<B, 1024, 0, st>>>(g0, coef, u0, C, per, norm);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_combine_losses_norm(const float* lv, const float* lf, const float* norm, int B, float gp_weight, float* out4, cudaStream_t st) {
    combine_losses_norm_kernel<<<1, 32, 0, st>>>(lv, lf, norm, B, gp_weight, out4);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_score_losses(const float* score, int B, int nloss, float s0, float s1, float* loss, cudaStream_t st) {
    score_losses_kernel<<<1, 64, 0, st>>>(score, B, nloss, s0, s1, loss);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_gp_loss(const float* norm, int B, float* out, cudaStream_t st) {
    gp_loss_kernel<<<1, 256, 0, st>>>(norm, B, out);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_extract_channel0(const float* x, float* out, long long n, int C, cudaStream_t st) {
    if (!n) return 0;
    extract_channel0_kernel<<<EW_GRID(n)>>>(x, out, n, C);
    RDG_LAUNCH_CHECK();
    return 0;
}
int ew_combine_losses(const float* lv, const float* lf, const float* lgp, float gp_weight, float* out4, cudaStream_t st) {
    combine_losses_kernel<<<1, 1, 0, st>>>(lv, lf, lgp, gp_weight, out4);
    RDG_LAUNCH_CHECK();
    return 0;
}
