// Training-path implicit GEMMs on the 5th-generation tensor cores (tcg_gemm.cu): every contraction of the cWGAN-GP step
// (gan_train_cwgangp_pixelnorm.py:272-357 forward, :230-244 gradient penalty, :387-408 gradients) as tcgen05.mma kind::tf32
// with FP32 accumulators in TMEM.  Operands are the FP32 training tensors themselves (rounded to tf32 on the way into shared
// memory), so the tensor-core mode shares every buffer, mask and pre-activation with the FP32 SIMT parity mode.
// Same argument meaning as the simt_* / folded_* primitives of rdg_common.cuh.
#pragma once
#include "rdg_common.cuh"

#define RDG_TCG_E_SHAPE (-12)   // geometry the tensor-core training kernels do not cover (callers fall back to nothing: it is an error)

// [nblk][R][C] -> [nblk][C][R]
// Split-K partial slices.  A step phase lends the row GEMMs one grow-only buffer for the duration of a TcgArenaScope: all split
// launches of a phase run on one stream and each use of the buffer ends with that launch's reduction kernel, so the next launch
// may overwrite it.  Without a scope (or on another stream than the phase's first split launch) a launch brackets itself with a
// stream-ordered allocation + free -- two extra nodes per launch on the dependency chain of a captured step.
struct TcgArena { float* buf = nullptr; size_t bytes = 0; cudaStream_t stream = nullptr; bool bound = false; };
struct TcgArenaScope {
    explicit TcgArenaScope(TcgArena* a);
    ~TcgArenaScope();
    TcgArena* prev;
};
int tcg_transpose_blocks(const float* src, float* dst, int nblk, int R, int C, cudaStream_t st);
// up to four of those in one launch
int tcg_transpose_blocks_batch(int n, const float* const* src, float* const* dst, const int* nblk, const int* R, const int* C, cudaStream_t st);

// act: ACT_NONE / ACT_LRELU, or ACT_LRELU_BWD with `pre` READ as the pre-activation whose LeakyReLU derivative multiplies the
// result (the linear second-order pass of the gradient penalty: u_l = conv_l(u_{l-1}) * LeakyReLU'(a_l) * mask).
// y = act(conv(x, w) + bias) [* mask * mask_scale]; wT = the kernel with its last two axes swapped: (KT,KH,KW,Co,Ci).
// g.up must be 0 (the generator's upsampled convs go through the folded entry points below).
// precise: 3xTF32 (operands split into tf32 high + low parts, three MMAs per k step: FP32-grade products).  Forward passes use it:
// LeakyReLU's derivative is discontinuous, so a pre-activation that changes sign under tf32 rounding changes every upstream
// gradient; backward passes are linear in their operands and run single-pass tf32.
int tcg_conv_fwd(const float* x, const float* wT, const float* bias, float* y, const ConvGeom& g, int act, const float* mask,
                 float mask_scale, cudaStream_t st, float* pre = nullptr, int precise = 0);
// few input channels (Ci <= 4: the critic's first conv, gan_train_cwgangp_pixelnorm.py:286-287): K = taps * Ci, gathered per element.
// wTp = tcg_pack_smallci_weights(w): [Co][Kpad], Kpad = tcg_smallci_kpad(taps, Ci), zero beyond taps * Ci.
int tcg_smallci_kpad(int taps, int Ci);
int tcg_pack_smallci_weights(const float* w, float* wTp, int taps, int Ci, int Co, cudaStream_t st);
int tcg_conv_fwd_smallci(const float* x, const float* wTp, const float* bias, float* y, const ConvGeom& g, int act, const float* mask,
                         float mask_scale, cudaStream_t st, float* pre = nullptr, int precise = 0);
int tcg_conv_bwd_filter_smallci(const float* x, const float* dy, float* dw, const ConvGeom& g, cudaStream_t st);
// scoring mode: first critic conv with the sample / condition concat fused into the gather, LeakyReLU, 16-bit output (half_kind:
// RDG_HALF_BF16 / RDG_HALF_FP16 of gen_tc.h)
// wTq (optional): tcg_pack_smallci_weights_chmajor image [Co][Ci * 32] for the sample-resident kernel used at nd = 16
int tcg_pack_smallci_weights_chmajor(const float* w, float* wTq, int taps, int Ci, int Co, cudaStream_t st);
int tcg_critic_first_conv16(int half_kind, const float* sample, const float* cond, const float* wTp, const float* wTq, const float* bias,
                            void* out16, const ConvGeom& g, cudaStream_t st);
// dx = conv_transpose(dy, w); w in the Keras layout (KT,KH,KW,Ci,Co).  stride 1 or 2, g.up == 0.
// pre_in != null: the LeakyReLU (+ dropout) backward of the layer below is fused into the epilogue: dx = conv_transpose(dy, w) *
// LeakyReLU'(pre_in) [* mask * mask_scale], i.e. the cotangent of that layer's PRE-activation (pre_in, mask: shaped like dx).
int tcg_conv_bwd_data(const float* dy, const float* w, float* dx, const ConvGeom& g, cudaStream_t st, const float* pre_in = nullptr,
                      const float* mask = nullptr, float mask_scale = 1.f);
// dw += sum_{b,pos} x (x) dy (accumulates with atomics; caller zeroes).  No bias gradient (use simt_colsum).
int tcg_conv_bwd_filter(const float* x, const float* dy, float* dw, const ConvGeom& g, cudaStream_t st);

// Upsample-folded forms of UpSampling3D(2) + Conv3D(3^3,'same'); g = the layer's geometry (g.up == 1).
// wfT: [8 phases][8 taps][Co][Ci] (tcg_transpose_blocks of folded_pack_f32's output); wf: [8][8][Ci][Co].
// y[B,2T,2H,2W,Co] = conv3(upsample2(x)) + bias
int tcg_folded_fwd(const float* x, const float* wfT, const float* bias, float* y, const ConvGeom& g, cudaStream_t st, int precise = 0);
// dx[B,T,H,W,Ci] = gradient w.r.t. the LOW-RES input (upsample backward absorbed); dy interleaved [B,2T,2H,2W,Co]
int tcg_folded_bwd_data(const float* dy, const float* wf, float* dx, const ConvGeom& g, cudaStream_t st);
// dwf[8][8][Ci][Co] += folded filter gradients (atomics; caller zeroes, then un-folds with folded_unfold_grad)
int tcg_folded_bwd_filter(const float* x, const float* dy, float* dwf, const ConvGeom& g, cudaStream_t st);
// dw[27][Ci][Co] += unfold(dwf)   (simt_folded.cu)
int folded_unfold_grad(const float* dwf, float* dw, int Ci, int Co, cudaStream_t st);
