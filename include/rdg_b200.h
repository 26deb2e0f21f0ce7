/* rdg_b200.h -- C ABI of the B200-native RainDisaggGAN hot path.
 *
 * Drop-in boundary for the ONE call boundary the reference has on this path: the Keras
 * model calls it makes into TensorFlow.  Every entry point cites the reference interface
 * it replaces (file:line under the reference repo).  Plain pointers and sizes only; no
 * torch / Python types.  All "dev" pointers are CUDA device pointers on the context's
 * device, "host" pointers are ordinary host memory.  Calls taking a stream are
 * stream-ordered and non-blocking unless stated otherwise; stream is a cudaStream_t
 * passed as void* (NULL = legacy default stream).
 *
 * Return value: 0 = ok, >0 = cudaError_t, <0 = RDG_E_* below.  rdg_last_error() returns a
 * thread-local message for the last non-zero return.
 *
 * Layouts (Keras channels-last, row-major):
 *   latent  [B,100] f32          cond   [Bc,nd,nd,ncond] f32 (already / norm_scale)
 *   sample  [B,24,nd,nd,1] f32   scores [B] f32
 *   Keras weights: Dense kernel (in,out); Conv3D kernel (kt,kh,kw,Cin,Cout); bias (Cout)
 */
#ifndef RDG_B200_H
#define RDG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDG_E_BADARG   (-1)
#define RDG_E_NOWEIGHT (-2)
#define RDG_E_NONFINITE (-3)  /* mirrors tf.debugging.check_numerics, gan_train_cwgangp_pixelnorm.py:349-350 */
#define RDG_E_NODEVICE (-4)
#define RDG_E_NOMEM    (-5)

/* arithmetic modes of the generator forward */
#define RDG_MODE_FP32 0   /* FP32 SIMT path: <=1e-5 relative vs the FP64 oracle              */
#define RDG_MODE_BF16 1   /* tcgen05 kind::f16, bf16 operands, f32 accumulate in TMEM        */
#define RDG_MODE_FP16 2   /* tcgen05 kind::f16, fp16 operands, f32 accumulate in TMEM        */

/* output scaling of the generator forward */
#define RDG_OUT_FRACTION 0  /* gen.predict: fractions, sum over the 24 hours == 1 (gan_train...py:347)      */
#define RDG_OUT_MM       1  /* fraction * cond * norm_scale, raindisagg_gan_pretrained.py:62-64 fused on-chip */

typedef struct rdg_ctx rdg_ctx;

const char* rdg_last_error(void);
int  rdg_version(void);

/* One context per (device, domain size). nd: 16 (gan_train_cwgangp_pixelnorm.py:51) or 64
 * (alternative_domains/...largedomain.py:59); ncond: condition channels (1; 2/3 for the
 * revision1 variants).  max_chunk = largest number of samples processed per internal pass
 * (workspace is sized from it; 0 = default). */
int  rdg_ctx_create(rdg_ctx** out, int device, int nd, int ncond, int max_chunk);
void rdg_ctx_destroy(rdg_ctx* ctx);
int  rdg_ctx_info(const rdg_ctx* ctx, int* nd, int* ncond, int* max_chunk, int* sm_count);
size_t rdg_ctx_workspace_bytes(const rdg_ctx* ctx);

/* Replaces tf.keras.models.load_model(...).get_weights() -> device
 * (raindisagg_gan_pretrained.py:43-45; gan_train_cwgangp_pixelnorm.py:361).  `tensors` are the
 * 10 generator arrays in Keras order (Dense k,b; Conv3D k,b x4), host f32.  Packs the FP32
 * copies, and the upsample-folded + swizzled 16-bit operand tiles for both tensor modes. */
int  rdg_generator_set_weights(rdg_ctx* ctx, const float* const* tensors, const size_t* sizes, int n);
int  rdg_generator_get_weights(rdg_ctx* ctx, float* const* tensors, const size_t* sizes, int n);

/* Replaces gen.predict([latent, cond_batch]) (raindisagg_gan_pretrained.py:60;
 * generate_and_evaluate_crps.py:185).  latent_dev [B,100]; cond_dev [ceil(B/scen_per_cond),
 * nd,nd,ncond] -- sample b uses condition b / scen_per_cond (scen_per_cond=1: one condition
 * per sample, as Keras).  out_dev [B,24,nd,nd] f32.  nonfinite_flag_dev (optional, int32)
 * is set !=0 if any output is NaN/Inf (check_numerics). */
int  rdg_generator_forward(rdg_ctx* ctx, const float* latent_dev, const float* cond_dev,
                           int scen_per_cond, float* out_dev, int B, int mode, int out_kind,
                           float norm_scale, int* nonfinite_flag_dev, void* stream);

/* Host-buffer end-to-end variant of the same call: chunks the batch and overlaps H2D of the next chunk / compute / D2H of the
 * previous one on three streams of the context.  Pinned (cudaHostAlloc / cudaHostRegister) buffers are copied from / to
 * directly; a PAGEABLE result buffer (a plain numpy array) larger than one chunk goes through an internal ring of three pinned
 * staging slots that host callbacks on a fourth stream drain into the caller's buffer, so the copies still overlap.
 * Blocking.  Runs on the context's own streams: if generator weights were changed on a caller stream the call first waits for
 * the device.  Returns RDG_E_NONFINITE like check_numerics would raise. */
int  rdg_generate_host(rdg_ctx* ctx, const float* latent_host, const float* cond_host,
                       int scen_per_cond, float* out_host, long long B, int mode, int out_kind,
                       float norm_scale);

/* Ensemble statistics on the device -- the host reductions that follow gen.predict in the
 * reference: area means np.mean(generated * cond * norm_scale, (2, 3)) (generate_and_evaluate.py:412-415, 533-535)
 * and properscoring.crps_ensemble(real_precip, generated, axis=0) with its area mean
 * (generate_and_evaluate_crps.py:189-191).  fields_dev [n_cond*scen_per_cond,24,nd,nd] (members of a condition
 * contiguous), obs_dev [n_cond,24,nd,nd] or NULL.  Outputs (each optional): area_mean_dev [n_cond*scen_per_cond,24],
 * crps_dev [n_cond,24,nd,nd], crps_area_mean_dev [n_cond,24]. */
int  rdg_ensemble_stats(rdg_ctx* ctx, const float* fields_dev, int n_cond, int scen_per_cond,
                        const float* obs_dev, float* area_mean_dev, float* crps_dev,
                        float* crps_area_mean_dev, void* stream);
/* rdg_generate_host + rdg_ensemble_stats fused per chunk: the fields never leave the GPU, only the statistics are
 * copied back (the whole CRPS loop of generate_and_evaluate_crps.py:177-191 in one call).  obs_host [n_cond,24,nd,nd]
 * mm/h or NULL; area_mean_host [B,24] or NULL; crps_area_mean_host [n_cond,24] or NULL.  B = n_cond * scen_per_cond. */
int  rdg_generate_stats_host(rdg_ctx* ctx, const float* latent_host, const float* cond_host, int scen_per_cond,
                             const float* obs_host, long long B, int mode, int out_kind, float norm_scale,
                             float* area_mean_host, float* crps_area_mean_host);

/* Device-side batch sampler: the preprocessing of generate_real_samples / generate_latent_points
 * (gan_train_cwgangp_pixelnorm.py:143-174, 177-193) on a radar array resident in HBM.  data_dev [n_days,24,ny,nx] f32 mm/h;
 * idx_dev [n,3] int32 rows (tidx, yidx, xidx) = indices_all[ixs] (:148); batch_dev [n,24,nd,nd] fractions of the daily sum
 * (NULL for generate_latent_points, which only needs the condition); cond_dev [n,nd,nd] = daily sum / norm_scale.
 * flag_dev (optional int32): bit 0 = a fraction is NaN or outside [0,1] (the reference's asserts :167-170), bit 1 = an
 * index outside the array (:133-136). */
int  rdg_sample_windows(const float* data_dev, int n_days, int ny, int nx, const int* idx_dev, int n, int nd,
                        float norm_scale, float* batch_dev, float* cond_dev, int* flag_dev, void* stream);

/* Device-side N(0,1) latent (counter-based Philox4x32-10 + Box-Muller), for throughput runs
 * where the reference would call np.random.normal (raindisagg_gan_pretrained.py:56). */
int  rdg_fill_normal(float* dst_dev, long long n, uint64_t seed, uint64_t offset, void* stream);

/* ---- critic (create_discriminator, gan_train_cwgangp_pixelnorm.py:272-309) ---- */
int  rdg_critic_set_weights(rdg_ctx* ctx, const float* const* tensors, const size_t* sizes, int n);
int  rdg_critic_get_weights(rdg_ctx* ctx, float* const* tensors, const size_t* sizes, int n);
/* critic([sample, cond]) -> score[B].  masks_dev: NULL (inference) or 4 device pointers to
 * {0,1} f32 dropout keep-masks shaped like each conv output (training mode, A9). */
int  rdg_critic_forward(rdg_ctx* ctx, const float* sample_dev, const float* cond_dev,
                        const float* const* masks_dev, float* score_dev, int B, void* stream);

/* critic([sample, cond]) in the 16-bit tensor-core scoring mode (mode = RDG_MODE_FP16 / RDG_MODE_BF16; inference, Dropout is
 * identity): the stride-2 convs D2..D4 (gan_train_cwgangp_pixelnorm.py:291-301) run as tcgen05 implicit GEMMs fed by TMA boxes
 * with element stride 2, the 2-channel first conv (:286-288) and the final Dense (:303-304) on CUDA cores. */
int  rdg_critic_forward_tc(rdg_ctx* ctx, const float* sample_dev, const float* cond_dev, float* score_dev, int B, int mode,
                           void* stream);

/* ---- training steps (gan_train_cwgangp_pixelnorm.py:365-408, 468-482) ----
 * Evaluate the losses and the gradient of the step's loss w.r.t. the trainable net, leaving
 * the gradients in the context's gradient buffers (for an allreduce between ranks), then
 * rdg_adam_apply() performs the Keras-Adam update (:385) with the shared step counter. */
int  rdg_critic_step_grads(rdg_ctx* ctx, const float* x_real_dev, const float* cond_dev,
                           const float* latent_dev, const float* alpha_dev,
                           const float* const* masks_fake, const float* const* masks_real,
                           const float* const* masks_hat, int B, int gen_mode,
                           float* losses4_dev, void* stream);
int  rdg_generator_step_grads(rdg_ctx* ctx, const float* latent_dev, const float* cond_dev,
                              const float* const* masks, int B, float* loss_dev, void* stream);
/* critic_model.predict's third output (gan_train_cwgangp_pixelnorm.py:461, :238-241): grad[b] = d D(sample_b, cond_b) / d sample_b
 * [B,24,nd,nd] (FP32, no dropout), optionally the scores [B] as well; GradientPenalty = sqrt(sum(grad^2)) - 1 per sample. */
int  rdg_critic_input_grad(rdg_ctx* ctx, const float* sample_dev, const float* cond_dev, float* score_dev,
                           float* grad_dev, int B, void* stream);
/* which: 0 = generator, 1 = critic.  Returns device pointer + element count of the flat
 * FP32 gradient / parameter buffers (Keras tensor order, concatenated). */
int  rdg_grad_buffer(rdg_ctx* ctx, int which, float** grads_dev, size_t* n);
int  rdg_param_buffer(rdg_ctx* ctx, int which, float** params_dev, size_t* n);
/* tf.optimizers.Adam(lr, beta_1, beta_2) OptimizerV2 semantics; step_t is the
 * already-incremented shared counter; grad_scale multiplies the gradient first (1/world). */
int  rdg_adam_apply(rdg_ctx* ctx, int which, float lr, float beta1, float beta2, float eps,
                    long long step_t, float grad_scale, void* stream);
int  rdg_adam_reset(rdg_ctx* ctx, int which);
/* Device pointers to the Adam moment buffers m, v (flat, same layout as rdg_param_buffer) -- optimizer-state checkpoints
 * (the reference saves weights only, gan_train_cwgangp_pixelnorm.py:520-521, and cannot resume). */
int  rdg_adam_buffers(rdg_ctx* ctx, int which, float** m_dev, float** v_dev, size_t* n);

/* ---- tensor-core training mode and replayable steps ----
 * rdg_set_train_mode(ctx, 1): the two step evaluations above run every wide contraction (critic convs forward / transposed /
 * filter gradients, the gradient-penalty second-order pass :230-244, the generator's upsample-folded convs and Dense, forward and
 * backward :387-408) as tcgen05.mma kind::tf32 implicit GEMMs with FP32 accumulation in TMEM (gradients <= 1e-2 relative L2 of
 * the FP64 oracle; measured ~1e-3).  Mode 0 (default) is the FP32 SIMT parity mode (<= 2e-5). */
int  rdg_set_train_mode(rdg_ctx* ctx, int mode);
/* The same evaluations with the step's random inputs drawn on the device (Philox: key = seed, counter = element / stream / a step
 * counter that lives in device memory): latent ~ N(0,1) (:470, :177-193), alpha ~ U[0,1) (:223), Dropout(0.25) keep masks
 * (:289-301; dropout = 0 disables them).  No host value enters the launch arguments, so the calls of a whole 5 + 1 iteration
 * (plus rdg_adam_apply_dev and the gradient all-reduce) can be captured in one CUDA graph and replayed.  Need train mode 1.
 * phases: 3 = the whole evaluation; 1 = only the part that does not read the critic's weights (random draws, generator forward,
 * interpolation, critic inputs), 2 = the rest.  Issuing 1 and 2 separately lets the head of a step run while another stream
 * still exchanges / applies the previous step's critic gradients.  rdg_critic_step_dev: phases | RDG_STEP_SLOT1 runs the call
 * on the context's SECOND set of step buffers (workspace + random inputs), both phases of one step with the same choice: with
 * alternating sets, phase 1 of step k+1 (frozen generator forward, draws) can run on another stream next to phase 2 of step k. */
#define RDG_STEP_SLOT1 16
int  rdg_critic_step_dev(rdg_ctx* ctx, const float* x_real_dev, const float* cond_dev, int B, int gen_mode,
                         unsigned long long seed, int dropout, float* losses4_dev, int phases, void* stream);
int  rdg_generator_step_dev(rdg_ctx* ctx, const float* cond_dev, int B, unsigned long long seed, int dropout,
                            float* loss_dev, int phases, void* stream);
/* ---- data-parallel gradient exchange over NVLink peer memory (one node, one process per GPU; csrc/dp_peer.cu) ----
 * Replaces the NCCL all-reduce a data-parallel run of the reference's train() (:463-491) would issue before each optimizer
 * update (:385, :412) by the library's own kernels, so that the exchange can be captured in the CUDA graph of an iteration.
 * rdg_peer_export: CUDA IPC handles (rdg_peer_handle_bytes() bytes) of this context's two gradient buffers and its flag array;
 * gather them from all ranks (any host-side transport) and pass the concatenation, in rank order, to rdg_peer_connect on every
 * rank.  rdg_peer_allreduce(ctx, which, stream): in-place SUM over the ranks of the gradient buffer (`which` as in
 * rdg_grad_buffer), stream-ordered: barrier, two-shot reduce (rank r sums the r-th slice in rank order and stores it to every
 * rank), barrier.  Every rank must make the same sequence of calls.  The 1/world scale belongs in rdg_adam_apply*'s grad_scale.
 * rdg_peer_status: world size (0 = not connected) and whether a barrier ever timed out (60 s) waiting for a peer
 * (the sums of that exchange are then not the global sums: treat it as a failed run). */
int  rdg_peer_handle_bytes(void);
int  rdg_peer_export(rdg_ctx* ctx, unsigned char* handles);
int  rdg_peer_connect(rdg_ctx* ctx, int world, int rank, const unsigned char* all_handles);
int  rdg_peer_disconnect(rdg_ctx* ctx);
int  rdg_peer_allreduce(rdg_ctx* ctx, int which, void* stream);
int  rdg_peer_status(rdg_ctx* ctx, int* world, int* timeouts);
/* rdg_adam_apply with the shared step counter kept (and incremented) in device memory, followed by the refresh of the derived
 * weight images the next step reads; gen_mode = the mode of the frozen generator forward in the critic step. */
int  rdg_adam_apply_dev(rdg_ctx* ctx, int which, float lr, float beta1, float beta2, float eps, float grad_scale,
                        int gen_mode, void* stream);
/* set = 0: read, set = 1: write the device-resident counters: the shared Adam step and the two Philox step counters rng_ctr2[0]
 * (critic steps) / rng_ctr2[1] (generator steps: phase 1 of a generator step may run next to the critic steps of its iteration, so
 * the two kinds of step count separately); synchronises the device. */
int  rdg_train_state(rdg_ctx* ctx, int set, long long* adam_t, unsigned long long* rng_ctr2);

/* ---- building blocks exposed for tests (same kernels the calls above use) ---- */
/* y = x / sqrt(mean_c(x^2) + 1e-8), optional LeakyReLU(0.2) (gan_train...py:255-266, :333) */
int  rdg_pixelnorm(const float* x_dev, float* y_dev, long long rows, int C, int lrelu, void* stream);
/* softmax over the hour axis of logits [B,24,P] (gan_train...py:347) */
int  rdg_softmax_hours(const float* logits_dev, float* out_dev, long long B, int P, void* stream);

/* The FP32 SIMT Conv3D primitives behind the critic / training path (Keras Conv3D semantics,
 * gan_train_cwgangp_pixelnorm.py:286-301, :331-345).  geom17 = {B, Ti,Hi,Wi,Ci, To,Ho,Wo,Co, KT,KH,KW, stride,
 * pad_before_t,h,w, upsample_input}.  op 0: out = act(conv(a, w=b) + bias); op 1: out = d/d(input) given
 * dy=a, w=b (w.r.t. the upsampled input if upsample_input); op 2: out += d/dw given x=a, dy=b, out2 += d/dbias.
 * op 10 / 11 / 12: the same three on the tensor cores (tcgen05 kind::tf32, the training mode's kernels); with upsample_input
 * they take the upsample-folded forms and op 11 returns the gradient w.r.t. the LOW-RES input. */
int  rdg_conv3d(int op, const int* geom17, const float* a, const float* b, const float* bias, float* out,
                float* out2, int act, void* stream);

/* instrumentation: per-layer CUDA-event timing on the launching stream (layer ids: 0 input concat,
 * 1 dense, 2 f32->16-bit, 3..5 upsampled convs, 6 output conv + softmax, 7 pixelnorm in FP32 mode)
 * and a count of kernels launched by this context. */
int  rdg_profile_enable(rdg_ctx* ctx, int on);
int  rdg_profile_collect(rdg_ctx* ctx, double* ms_sum8, long long* launches8, long long* units8);
long long rdg_launch_count(const rdg_ctx* ctx);

/* one tensor-core generator layer (0..2) in isolation: UpSampling3D + Conv3D + PixelNorm + LeakyReLU
 * (gan_train_cwgangp_pixelnorm.py:330-343); x f32 [B,T,H,W,Cin] is rounded to the mode's 16-bit type,
 * y f32 [B,2T,2H,2W,Cout] is the widened 16-bit result. */
int  rdg_tc_layer(rdg_ctx* ctx, int layer, int mode, const float* x_dev, float* y_dev, int B, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RDG_B200_H */
