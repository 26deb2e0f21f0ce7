// Shared declarations for the RainDisaggGAN sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define RDG_NHOURS 24
#define RDG_LATENT 100

// Geometry of one 3-D convolution over channels-last tensors (B,T,H,W,C).
// Physical input dims (Ti,Hi,Wi); when `up` is set the kernel reads the input through a
// nearest x2 upsample (logical dims 2Ti,2Hi,2Wi) -- UpSampling3D fused into the gather.
struct ConvGeom {
    int B;
    int Ti, Hi, Wi, Ci;
    int To, Ho, Wo, Co;
    int KT, KH, KW;
    int stride;
    int pt, ph, pw;   // zero padding BEFORE (TF 'same' puts the odd one after)
    int up;
};

enum { ACT_NONE = 0, ACT_LRELU = 1, ACT_ACCUM = 2 /* backward-data only: dst += result */,
       ACT_LRELU_BWD = 3 /* tcg kernels: result *= LeakyReLU'(pre) (pre is READ: the pre-activation the gradient flows into) [* mask] */ };

void rdg_set_error(const char* fmt, ...);

#define RDG_CUDA(call)                                                                     \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            rdg_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return (int)e_;                                                                \
        }                                                                                  \
    } while (0)

#define RDG_LAUNCH_CHECK()                                                                 \
    do {                                                                                   \
        cudaError_t e_ = cudaGetLastError();                                               \
        if (e_ != cudaSuccess) {                                                           \
            rdg_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            return (int)e_;                                                                \
        }                                                                                  \
    } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- SIMT FP32 primitives (simt_conv.cu) ----
// y = act(conv(x, w) + bias) [* mask * mask_scale]; w is Keras (KT,KH,KW,Ci,Co) f32.
int simt_conv_fwd(const float* x, const float* w, const float* bias, float* y, const ConvGeom& g,
                  int act, const float* mask, float mask_scale, cudaStream_t st, float* pre = nullptr);
// out[c] += sum_r x[r][c]
int simt_colsum(const float* x, float* out, long long rows, int C, cudaStream_t st);
// dx = conv_transpose(dy, w): gradient w.r.t. the (logical, i.e. upsampled if g.up) input.
int simt_conv_bwd_data(const float* dy, const float* w, float* dx, const ConvGeom& g, cudaStream_t st, int accumulate = 0);
int simt_conv_bwd_data_ch0(const float* dy, const float* w, float* dx, const ConvGeom& g, cudaStream_t st,
                           float* scratch = nullptr);   // input channel 0 only; scratch: ntaps * B*To*Ho*Wo floats (tap-product form)

// ---- upsample-folded FP32 forms of UpSampling3D(2) + Conv3D(3^3,'same') (simt_folded.cu) ----
// wf: [8 phases][2,2,2,Ci,Co] folded kernels (f32 sums of the 3^3 kernel's taps); g = the layer's geometry (g.up == 1).
size_t folded_weight_elems(int Ci, int Co);
int folded_pack_f32(const float* k, float* wf, int Ci, int Co, cudaStream_t st);
// y[B,2T,2H,2W,Co] = conv3(upsample2(x)) + bias; scratch: B*2T*2H*2W*Co floats (phase-major staging)
int folded_conv_fwd(const float* x, const float* wf, const float* bias, float* y, float* scratch, const ConvGeom& g, cudaStream_t st);
// dyp (out): phase-major copy of dy [8][B,T,H,W,Co]; dx[B,T,H,W,Ci] = gradient w.r.t. the LOW-RES input (upsample backward included)
int folded_conv_bwd_data(const float* dy, const float* wf, float* dx, float* dyp, const ConvGeom& g, cudaStream_t st);
// the two halves of folded_conv_bwd_data, for callers that overlap the filter gradient with the second one
int folded_deinterleave(const float* dy, float* dyp, const ConvGeom& g, cudaStream_t st);
int folded_conv_bwd_data_phases(const float* dyp, const float* wf, float* dx, const ConvGeom& g, cudaStream_t st);
// dw[27,Ci,Co] += unfold(sum_p bwd_filter(x, dyp[p])); db[Co] += colsum(dy).  dwf: scratch of folded_weight_elems floats.
int folded_conv_bwd_filter(const float* x, const float* dy, const float* dyp, float* dwf, float* dw, float* db, const ConvGeom& g, cudaStream_t st);
// dw += sum_b,pos x (x) dy ; db += sum dy (db may be null). Accumulates (caller zeroes).
int simt_conv_bwd_filter(const float* x, const float* dy, float* dw, float* db, const ConvGeom& g,
                         cudaStream_t st);

// ---- elementwise / reductions (elementwise.cu) ----
int ew_assemble_gen_input(const float* latent, const float* cond, int scen_per_cond, int b_off, float* x0,
                          int B, int ncondflat, cudaStream_t st);
int ew_pixelnorm(const float* x, float* y, long long rows, int C, int lrelu, cudaStream_t st);
int ew_pixelnorm_lrelu_bwd(const float* x_pre, const float* dy, float* dx, long long rows, int C,
                           cudaStream_t st);
int ew_softmax_hours(const float* logits, float* out, long long B, int P, const float* cond,
                     int scen_per_cond, int ncond, float scale, int out_mm, int* nonfinite,
                     cudaStream_t st);
int ew_softmax_hours_bwd(const float* y, const float* dy, float* dlogits, long long B, int P,
                         cudaStream_t st);
int ew_lrelu_bwd(const float* pre, const float* dy, float* dx, long long n, const float* mask,
                 float mask_scale, cudaStream_t st);
int ew_upsample_pool(const float* d_up, float* d_lo, int B, int T, int H, int W, int C, cudaStream_t st);
int ew_critic_input(const float* sample, const float* cond, float* x, int B, int nd, int ncond,
                    cudaStream_t st);
int ew_fill_normal(float* dst, long long n, uint64_t seed, uint64_t offset, cudaStream_t st);
int ew_adam(float* p, const float* g, float* m, float* v, long long n, float lr_t, float beta1,
            float beta2, float eps, float grad_scale, cudaStream_t st);

// device-resident training state (replayable CUDA graphs): Philox step counter, shared Adam step counter, bias-corrected rate
struct RdgTrainState { unsigned long long rng_ctr[2]; long long adam_t; float lr_t; float pad_; unsigned int done[2]; unsigned int pad2_[2]; };   // rng_ctr: [0] critic steps, [1] generator steps
int ew_train_tick(RdgTrainState* s, int which, cudaStream_t st);
// All random inputs of one step in ONE launch, counter advanced by the launch itself (the last block to finish increments it; every
// thread uses the old value + 1): buf[0 .. n_normal) ~ N(0,1) (stream 1), buf[off_u .. +n_u) ~ U[0,1) (stream 2), buf[off_m .. +n_m)
// Bernoulli(keep) (stream 3).  Same values as ew_train_tick + three ew_fill_random_dev calls.
int ew_step_random_dev(float* buf, long long n_normal, long long off_u, long long n_u, long long off_m, long long n_m, float keep, uint64_t seed,
                       RdgTrainState* s, int which, cudaStream_t st);     // which: 0 critic-step counter, 1 generator-step counter
// Keras-Adam with the step counter in *s (incremented here) instead of a host argument
int ew_adam_dev(float* p, const float* g, float* m, float* v, long long n, RdgTrainState* s, float lr, float beta1, float beta2,
                float eps, float grad_scale, cudaStream_t st);
// kind 0 N(0,1), 1 U[0,1), 2 Bernoulli(keep) 0/1; Philox key = seed, counter = (index, stream_id, s->rng_ctr[which])
int ew_fill_random_dev(float* dst, long long n, uint64_t seed, const RdgTrainState* s, int which, uint32_t stream_id, int kind, float keep,
                       cudaStream_t st);

// training-step helpers
int ew_tap_gather_logits(const float* P, const float* b4, float* logits, int B, int nd, cudaStream_t st);   // P [B,24,nd,nd,32]
int ew_tap_scatter_dlogits(const float* dl, float* Gd, int B, int nd, cudaStream_t st);                      // Gd [B,24,nd,nd,32]
int ew_pad_w4(const float* w4, float* w4p, cudaStream_t st);     // -> [32][64] then [64][32]
int ew_fill3(float* dst, int n, float a, float b, float c, cudaStream_t st);
// x3 [3B,24,nd,nd,1+ncond] = critic inputs [fake | real | alpha*real + (1-alpha)*fake] with the condition tiled over the hours
int ew_critic_inputs3(const float* fake, const float* real, const float* alpha, const float* cond, float* x3, int B, int nd, int ncond,
                      cudaStream_t st);
// critic forward tail + backward head: dscore[b] = cot3[b / B], loss[k] = sign2[k] * mean(score[kB .. kB+B)) for k < nloss,
// da4 = dscore x w5 * LeakyReLU'(a4) [* mask * mask_scale]   (nseg segments of B samples, K = flattened features)
int ew_critic_tail(const float* score, const float* w5, const float* a4, const float* mask, float mask_scale, int B, int nseg, int K,
                   const float* cot3, const float* sign2, int nloss, float* loss, float* dscore, float* da4, cudaStream_t st);
int ew_dense_score(const float* x, const float* w, const float* bias, float* score, int B, int K, cudaStream_t st);   // Flatten + Dense(1)
int ew_interp(const float* xr, const float* xf, const float* alpha, float* xhat, int B, long long per, cudaStream_t st);
int ew_fill(float* dst, long long n, float v, cudaStream_t st);
int ew_mean_scaled(const float* x, long long n, float scale, float* out, cudaStream_t st);   // out = scale * mean(x)
int ew_gp_norm(const float* g0, int C, int B, long long per, float* norm, cudaStream_t st);   // ||g0[b,...,0]||_2
int ew_gp_cotangent(const float* g0, const float* norm, float coef, float* u0, int C, int B, long long per, cudaStream_t st,
                    float* loss_out = nullptr);      // loss_out != null: also mean((norm - 1)^2) (saves the ew_gp_loss launch)
int ew_gp_loss(const float* norm, int B, float* out, cudaStream_t st);                       // mean((norm-1)^2)
int ew_extract_channel0(const float* x, float* out, long long n, int C, cudaStream_t st);
int ew_combine_losses(const float* lv, const float* lf, const float* lgp, float gp_weight, float* out4, cudaStream_t st);
// tensor-core training mode: norm + cotangent of the penalty in one launch (one block per sample), the penalty loss formed from
// norm[] when the losses are combined, and the Wasserstein loss terms as their own (off-chain) launch
int ew_gp_fused(const float* g0, float coef, float* u0, int C, int B, long long per, float* norm, cudaStream_t st);
int ew_combine_losses_norm(const float* lv, const float* lf, const float* norm, int B, float gp_weight, float* out4, cudaStream_t st);
int ew_score_losses(const float* score, int B, int nloss, float s0, float s1, float* loss, cudaStream_t st);
