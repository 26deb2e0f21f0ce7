// Host entry points of the tensor-core generator kernels (gen_tc.cu).
#pragma once
#include <cuda_runtime.h>

#define RDG_HALF_F32  0
#define RDG_HALF_BF16 1
#define RDG_HALF_FP16 2

#define RDG_TC_E_DRIVER (-10)
#define RDG_TC_E_SHAPE  (-11)

struct TcConvArgs {
    int B;              // samples in this launch
    int T, H, W;        // low-res input grid
    int Cin;
    int Hb, Bt;         // TMA box: Bt samples x Hb rows x W columns = 128 accumulator rows
    int n_tiles;
    int pass_split;     // generic kernel: CTAs (pairs) sharing the 8 / NPH passes of one tile (1 unless the launch is small)
    const void* wpack;  // [8 phases][8 taps][Cin/64][Cout rows x 64 k] 16-bit, rows pre-swizzled (128B)
    const float* bias;
    void* out;          // [B,2T,2H,2W,Cout] 16-bit (unused when the output conv is fused)
    const void* w4tile; // fused output conv: [32 taps x 64 ch] 16-bit swizzled tile (taps >= 27 zero)
    float* p_out;       // fused output conv: [B,2T,2H,2W,32] f32 per-tap partial products
    int* logit_out;     // planes kernel only: fused output conv summed on chip -> [B,2T,16,16] fixed-point logits (no bias)
    int sup_split;      // planes kernel: 1 = one work item per (sample-pair group, super-tile) (small launches)
    int* nonfinite;     // planes kernel, logit mode: set to 1 if any activation is Inf / NaN
};

// y[B,2T,2H,2W,Cout] = LeakyReLU(PixelNorm(conv3x3x3(upsample2(x[B,T,H,W,Cin])) + bias))
// If p_out != null (Cout == 64 only) the output Conv3D(64->1) tap products P are produced instead of y.
int tc_upconv_pixelnorm(int half_kind, const void* x, const void* wpack, const float* bias, void* y, const void* w4tile,
                        float* p_out, int B, int T, int H, int W, int Cin, int Cout, int sm_count, cudaStream_t st);
// out[B,24,nd,nd] = softmax_hours(b + sum_taps P[q+off(tap)][tap]) (* cond * scale), P [B,24,nd,nd,32] f32
int gather_softmax(const float* p, const float* b4, float* out, const float* cond, int B, int nd, int spc, int b_off,
                   int ncond, float scale, int out_mm, int* nonfinite, cudaStream_t st);
// out[B,24,nd,nd] = softmax_hours(conv3x3x3(x[B,24,nd,nd,64]) + b) (* cond * scale)
int conv_out_softmax(int in_kind, const void* x, const float* w4, const float* b4, float* out, const float* cond,
                     int B, int nd, int spc, int b_off, int ncond, float scale, int out_mm, int* nonfinite, cudaStream_t st);
int f32_to_half(int half_kind, const float* src, void* dst, long long n, cudaStream_t st);
int half_to_f32(int half_kind, const void* src, float* dst, long long n, cudaStream_t st);
// device-side packing of the master f32 weights into the tcgen05 operand images
int pack_folded_weights(int half_kind, const float* k, void* dst, int Cin, int Cout, cudaStream_t st);
int pack_w4_tile(int half_kind, const float* k4, void* dst, cudaStream_t st);

// Resident-plane kernel for the 128 -> 64 layer on an 8x8 low-res grid (ndomain 16), gen_tc_planes.cu.
// y (16-bit, [B,2T,16,16,64]) or, if p_out != null, the fused output-conv tap products P, or, if logit_out != null,
// the output-conv logits (without bias) as int32 fixed point with the scale pack_w4_tile stored behind the w4 tile.
int tc_upconv64_planes(int half_kind, const void* x, const void* wpack_planes, const float* bias, void* y, const void* w4tile,
                       float* p_out, int* logit_out, int* nonfinite, int B, int T, int sm_count, cudaStream_t st);
// in place: out[B,24,256] int32 fixed-point logits -> softmax over the hours of (v / scale + b) (* cond * norm_scale)
int softmax_fixed_inplace(float* out, const void* w4tile, const float* b4, const float* cond, int B, int spc, int b_off,
                          int ncond, float norm_scale, int out_mm, int* nonfinite, cudaStream_t st);
#define RDG_W4PACK_BYTES (32 * 64 * 2 + 16)   // swizzled tile + {fixed-point scale, 1/scale} (f32)
int pack_folded_weights_planes(int half_kind, const float* k, void* dst, cudaStream_t st);

// Fused Concatenate + Dense(3072) + LeakyReLU on tcgen05 (nd == 16, ncond == 1 only), gen_dense_tc.cu:
// out[B,3072] 16-bit = lrelu([latent | cond[(b_off+b)/spc]] @ W + bias); wpack from pack_dense_weights.
int tc_dense_lrelu(int half_kind, const float* latent, const float* cond, int spc, int b_off, const void* wpack,
                   const float* bias, void* out, int B, cudaStream_t st);
size_t tc_dense_pack_bytes();
int pack_dense_weights(int half_kind, const float* w, void* dst, cudaStream_t st);

// General Concatenate + Dense + LeakyReLU on tcgen05 (large domain, ncond > 1), gen_dense_big_tc.cu: out[B,N] 16-bit =
// lrelu([latent | cond[(b_off+b)/spc]] @ W + bias); wT from pack_dense_big_weights ([N][Kp] 16-bit), x16_scratch of
// tc_dense_big_input_bytes(B, K) bytes.
size_t tc_dense_big_pack_bytes(int K, int N);
size_t tc_dense_big_input_bytes(int B, int K);
int pack_dense_big_weights(int half_kind, const float* w, void* dst, int K, int N, cudaStream_t st);
int tc_dense_big_lrelu(int half_kind, const float* latent, const float* cond, int spc, int b_off, const void* wT, const float* bias,
                       void* x16_scratch, void* out, int B, int K, int N, cudaStream_t st);

// Critic forward on the tensor cores (critic_tc.cu): stride-2 Conv3D + LeakyReLU layers D2..D4 as tcgen05 implicit GEMMs
// (g = rdg_critic_conv_geom of the layer; x, y 16-bit channels-last), the 2-channel first conv and the final Dense on CUDA cores.
struct ConvGeom;
int tc_critic_conv(int half_kind, const void* x, const void* wpack, const float* bias, void* y, const ConvGeom& g, int sm_count, cudaStream_t st);
int pack_critic_weights(int half_kind, const float* k, void* dst, int Cin, int Cout, cudaStream_t st);
int critic_first_conv(int half_kind, const float* sample, const float* cond, const float* w, const float* bias, void* out, int B, int nd,
                      int ncond, const ConvGeom& g, cudaStream_t st);
int critic_dense_score(int half_kind, const void* h4, const float* w5, const float* b5, float* score, int B, int K, cudaStream_t st);
