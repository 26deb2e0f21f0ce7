// Ensemble statistics of generated scenarios on the device: the reductions the reference does on the
// host after every gen.predict call --
//   area means            np.mean(generated * cond * norm_scale, (2, 3))          generate_and_evaluate.py:412-415, 533-535
//   CRPS of the ensemble  properscoring.crps_ensemble(real_precip, generated, axis=0) and its area mean
//                                                                                  generate_and_evaluate_crps.py:189-191
// so that only [B,24] + [n_cond,24] floats leave the GPU instead of 24 KB per scenario.
// crps_ensemble (properscoring 0.1, equal weights) integrates (F_ens(x) - 1{x >= obs})^2 dx over the sorted members; for an
// empirical CDF that integral equals  mean_i |x_i - y|  -  (1/n^2) * sum_{i<j} |x_i - x_j|, which is what is evaluated here
// (no sort; O(n^2) per grid point in shared memory -- fine for the 100-member ensembles of config #2).
#include "rdg_common.cuh"
#include "ctx.h"
#include "../../include/rdg_b200.h"

namespace {

// out[r] = mean_p x[r][p]; one warp per row (row = (scenario, hour) or (condition, hour)), rows of npix floats
__global__ void __launch_bounds__(256) row_mean_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows, int npix) {
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    const float4* p = reinterpret_cast<const float4*>(x + r * npix);
    float s = 0.f;
    for (int i = lane; i < npix / 4; i += 32) { const float4 v = p[i]; s += (v.x + v.y) + (v.z + v.w); }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[r] = s / (float)npix;
}

// grid (n_cond, 24*npix/32), 256 threads: 32 consecutive (hour, pixel) columns x all spc members in shared memory.
__global__ void __launch_bounds__(256) crps_ensemble_kernel(const float* __restrict__ fields, const float* __restrict__ obs,
                                                            float* __restrict__ crps, int spc, int cols) {
    extern __shared__ float tile[];                       // [spc][32]
    __shared__ double red_pair[8][32], red_abs[8][32];
    const int cond = blockIdx.x, c0 = blockIdx.y * 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float* base = fields + (size_t)cond * spc * cols + c0 + lane;
    for (int m = w; m < spc; m += 8) tile[m * 32 + lane] = base[(size_t)m * cols];
    __syncthreads();
    const float y = obs[(size_t)cond * cols + c0 + lane];
    double pair = 0.0, ab = 0.0;
    for (int i = w; i < spc; i += 8) {
        const float xi = tile[i * 32 + lane];
        float s0 = 0.f, s1 = 0.f;
        int j = i + 1;
        for (; j + 1 < spc; j += 2) { s0 += fabsf(xi - tile[j * 32 + lane]); s1 += fabsf(xi - tile[(j + 1) * 32 + lane]); }
        if (j < spc) s0 += fabsf(xi - tile[j * 32 + lane]);
        pair += (double)(s0 + s1);
        ab += (double)fabsf(xi - y);
    }
    red_pair[w][lane] = pair; red_abs[w][lane] = ab;
    __syncthreads();
    if (w == 0) {
        for (int k = 1; k < 8; ++k) { pair += red_pair[k][lane]; ab += red_abs[k][lane]; }
        crps[(size_t)cond * cols + c0 + lane] = (float)(ab / spc - pair / ((double)spc * spc));
    }
}

// Large ensembles (spc > 160, e.g. the 1000 members of generate_and_evaluate_crps.py:161-162): bitonic sort of the members of
// each column in shared memory (padded to a power of two with +inf), then sum_{i<j} |x_i - x_j| = sum_k (2k - n + 1) x_(k):
// O(n log^2 n) instead of O(n^2) per grid point.
__global__ void __launch_bounds__(256) crps_ensemble_sort_kernel(const float* __restrict__ fields, const float* __restrict__ obs,
                                                                 float* __restrict__ crps, int spc, int npad, int cols) {
    extern __shared__ float tile[];                       // [npad][32]
    __shared__ double red_pair[8][32], red_abs[8][32];
    const int cond = blockIdx.x, c0 = blockIdx.y * 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float* base = fields + (size_t)cond * spc * cols + c0 + lane;
    for (int m = w; m < npad; m += 8) tile[m * 32 + lane] = m < spc ? base[(size_t)m * cols] : INFINITY;
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = w; t < npad / 2; t += 8) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                const float a = tile[i * 32 + lane], b = tile[l * 32 + lane];
                const bool asc = (i & k) == 0;
                if ((a > b) == asc) { tile[i * 32 + lane] = b; tile[l * 32 + lane] = a; }
            }
            __syncthreads();
        }
    const float y = obs[(size_t)cond * cols + c0 + lane];
    double pair = 0.0, ab = 0.0;
    for (int k = w; k < spc; k += 8) {
        const float x = tile[k * 32 + lane];
        pair += (double)(2 * k - spc + 1) * (double)x;
        ab += (double)fabsf(x - y);
    }
    red_pair[w][lane] = pair; red_abs[w][lane] = ab;
    __syncthreads();
    if (w == 0) {
        for (int k = 1; k < 8; ++k) { pair += red_pair[k][lane]; ab += red_abs[k][lane]; }
        crps[(size_t)cond * cols + c0 + lane] = (float)(ab / spc - pair / ((double)spc * spc));
    }
}

}  // namespace

static int stats_launch(rdg_ctx* c, const float* fields, int n_cond, int spc, const float* obs, float* area_mean, float* crps,
                        float* crps_amean, cudaStream_t st) {
    const int npix = c->nd * c->nd, cols = RDG_NHOURS * npix;
    if (area_mean) {
        const long long rows = (long long)n_cond * spc * RDG_NHOURS;
        row_mean_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(fields, area_mean, rows, npix);
        RDG_LAUNCH_CHECK();
        c->launches += 1;
    }
    if (obs && (crps || crps_amean)) {
        const size_t smem = (size_t)spc * 32 * sizeof(float);
        if (smem > 200 * 1024) { rdg_set_error("rdg_ensemble_stats: at most 1600 members per condition"); return RDG_E_BADARG; }
        float* tmp = nullptr;
        if (!crps) { RDG_CUDA(cudaMallocAsync(&tmp, (size_t)n_cond * cols * sizeof(float), st)); crps = tmp; }
        static bool attr_set = false;
        if (!attr_set) {
            RDG_CUDA(cudaFuncSetAttribute(crps_ensemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            RDG_CUDA(cudaFuncSetAttribute(crps_ensemble_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set = true;
        }
        int npad = 1;
        while (npad < spc) npad <<= 1;
        if (spc > 160 && (size_t)npad * 32 * sizeof(float) <= 200 * 1024)
            crps_ensemble_sort_kernel<<<dim3(n_cond, cols / 32), 256, (size_t)npad * 32 * sizeof(float), st>>>(fields, obs, crps, spc, npad, cols);
        else
            crps_ensemble_kernel<<<dim3(n_cond, cols / 32), 256, smem, st>>>(fields, obs, crps, spc, cols);
        RDG_LAUNCH_CHECK();
        c->launches += 1;
        if (crps_amean) {
            row_mean_kernel<<<ceil_div((long long)n_cond * RDG_NHOURS, 8), 256, 0, st>>>(crps, crps_amean, (long long)n_cond * RDG_NHOURS, npix);
            RDG_LAUNCH_CHECK();
            c->launches += 1;
        }
        if (tmp) RDG_CUDA(cudaFreeAsync(tmp, st));
    }
    return 0;
}

extern "C" int rdg_ensemble_stats(rdg_ctx* c, const float* fields_dev, int n_cond, int spc, const float* obs_dev,
                                  float* area_mean_dev, float* crps_dev, float* crps_area_mean_dev, void* stream) {
    if (!c || !fields_dev || n_cond < 0 || spc < 1) { rdg_set_error("rdg_ensemble_stats: bad arguments"); return RDG_E_BADARG; }
    if ((crps_dev || crps_area_mean_dev) && !obs_dev) { rdg_set_error("rdg_ensemble_stats: CRPS needs observations"); return RDG_E_BADARG; }
    if (n_cond == 0) return 0;
    RDG_CUDA(cudaSetDevice(c->device));
    return stats_launch(c, fields_dev, n_cond, spc, obs_dev, area_mean_dev, crps_dev, crps_area_mean_dev, (cudaStream_t)stream);
}

int rdg_stats_chunk(rdg_ctx* c, const float* fields, int n_cond, int spc, const float* obs, float* area_mean, float* crps_scratch,
                    float* crps_amean, cudaStream_t st) {
    return stats_launch(c, fields, n_cond, spc, obs, area_mean, crps_scratch, crps_amean, st);
}
