// Device-side batch sampler: the last preprocessing step of generate_real_samples /
// generate_latent_points (gan_train_cwgangp_pixelnorm.py:143-174, 177-193) on a radar array resident in HBM --
// window gather data[t, :, y:y+nd, x:x+nd] (view_as_windows, :151-152), daily sum over the 24 hours (:156),
// fraction normalisation batch / daily_sum (:159-160) and cond = daily_sum / norm_scale (:163).
// FP32, hours summed in order like np.sum(axis=1): bit-identical to the numpy statements.  HBM-bound gather:
// 24 KB read + 25 KB written per sample at nd = 16.
#include "rdg_common.cuh"
#include "../../include/rdg_b200.h"

namespace {

// one thread per (sample, pixel); consecutive threads = consecutive x: 64-byte runs per window row
__global__ void __launch_bounds__(256) sample_windows_kernel(const float* __restrict__ data, int n_days, int ny, int nx,
                                                             const int* __restrict__ idx, int n, int nd, float norm_scale,
                                                             float* __restrict__ batch, float* __restrict__ cond,
                                                             int* __restrict__ flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int npix = nd * nd;
    if (i >= (long long)n * npix) return;
    const int s = (int)(i / npix), p = (int)(i % npix), py = p / nd, px = p % nd;
    const int t = idx[3 * s], y = idx[3 * s + 1], x = idx[3 * s + 2];
    if (t < 0 || t >= n_days || y < 0 || y + nd > ny || x < 0 || x + nd > nx) {     // reference asserts :133-136
        if (flag) atomicOr(flag, 2);
        return;
    }
    const float* src = data + (((size_t)t * RDG_NHOURS) * ny + (y + py)) * nx + (x + px);
    float v[RDG_NHOURS];
    float sum = 0.f;
#pragma unroll
    for (int h = 0; h < RDG_NHOURS; ++h) { v[h] = src[(size_t)h * ny * nx]; sum += v[h]; }
    bool bad = !isfinite(sum);
    if (batch) {
        float* dst = batch + (size_t)s * RDG_NHOURS * npix + p;
#pragma unroll
        for (int h = 0; h < RDG_NHOURS; ++h) {
            const float f = v[h] / sum;                          // 0/0 -> NaN like numpy; the reference asserts on it (:167)
            bad |= !(f >= 0.f && f <= 1.f);                      // also catches NaN (:169-170)
            dst[(size_t)h * npix] = f;
        }
    }
    cond[(size_t)s * npix + p] = sum / norm_scale;
    if (bad && flag) atomicOr(flag, 1);
}

}  // namespace

extern "C" int rdg_sample_windows(const float* data_dev, int n_days, int ny, int nx, const int* idx_dev, int n, int nd,
                                  float norm_scale, float* batch_dev, float* cond_dev, int* flag_dev, void* stream) {
    if (n == 0) return 0;
    if (!data_dev || !idx_dev || !cond_dev || n < 0 || nd < 1 || nd > ny || nd > nx) { rdg_set_error("rdg_sample_windows: bad arguments"); return RDG_E_BADARG; }
    sample_windows_kernel<<<ceil_div((long long)n * nd * nd, 256), 256, 0, (cudaStream_t)stream>>>(
        data_dev, n_days, ny, nx, idx_dev, n, nd, norm_scale, batch_dev, cond_dev, flag_dev);
    RDG_LAUNCH_CHECK();
    return 0;
}
