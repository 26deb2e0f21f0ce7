// Context object behind the opaque rdg_ctx handle of include/rdg_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <vector>
#include "rdg_common.cuh"

#include "tcg.h"

struct rdg_ctx {
    int device = 0, nd = 16, ncond = 1, max_chunk = 0, sm_count = 0;
    // FP32 master parameters, Keras tensor order, concatenated (each tensor 16-byte aligned)
    float* g_params = nullptr; size_t g_off[10] = {}, g_size[10] = {}, g_total = 0;
    float* c_params = nullptr; size_t c_off[10] = {}, c_size[10] = {}, c_total = 0;
    bool gen_ready = false, critic_ready = false;
    int gen_stale_kinds = 0;        // 16-bit operand images older than the master weights: bit 0 = bf16, bit 1 = fp16
    // folded + swizzled 16-bit operand tiles of the three upsampled convs: [0]=bf16, [1]=fp16
    void* g_wpack[2][3] = {};
    void* g_wpack_planes[2] = {};   // 128 -> 64 layer in the stage order of the resident-plane kernel (nd == 16 only)
    bool conv3_planes = false;      // nd == 16 and not disabled by RDG_CONV3=tiles
    bool conv3_logits = true;       // planes kernel sums the output conv on chip (RDG_CONV3=planes_p: P through HBM)
    void* g_wpack_dense[2] = {};    // Dense kernel as tcgen05 B tiles (nd == 16, ncond == 1 only)
    bool dense_tc = false;
    void* g_wpack_dense_big[2] = {}; // general tensor-core Dense: 16-bit transpose [N][Kp] (everything dense_tc does not cover)
    bool dense_big_tc = false;
    float* g_wfold32[3] = {};       // FP32 upsample-folded kernels [8 phases][2,2,2,Ci,Co] of the three upsampled convs (simt_folded.cu)
    bool fold32_stale = true;
    void* c_wpack[2][3] = {};       // critic D2..D4 kernels as tcgen05 B tiles ([0] = bf16, [1] = fp16), critic_tc.cu
    bool critic_packed_stale = true;
    void* g_w4pack[2] = {};   // output conv as a [32 taps x 64 ch] swizzled 16-bit B tile
    // training state (allocated on first use)
    float* g_grads = nullptr; float* g_m = nullptr; float* g_v = nullptr;
    float* c_grads = nullptr; float* c_m = nullptr; float* c_v = nullptr;
    void* train_ws = nullptr; size_t train_ws_bytes = 0;
    void* train_ws1 = nullptr; size_t train_ws1_bytes = 0;     // second critic-step workspace + random inputs (RDG_STEP_SLOT1):
    float* rnd_buf1 = nullptr; size_t rnd1_cap = 0;            // lets phase 1 of step k+1 run next to phase 2 of step k
    // the generator step of the tensor-core mode has its own workspace and random-input buffer: its first phase (generator forward)
    // reads no critic state and runs on another stream next to the critic steps of the same iteration
    void* train_ws_gen = nullptr; size_t train_ws_gen_bytes = 0;
    float* rnd_buf_gen = nullptr; size_t rnd_gen_cap = 0;
    // tensor-core training mode (tcg_gemm.cu, train_tc.cu): 0 = FP32 SIMT (<= 1e-5 parity mode), 1 = tcgen05 kind::tf32
    int train_mode = 0;
    float* c_w1q_score = nullptr;                    // channel-major image [Co][Ci * 32] for the sample-resident first-conv kernel (nd = 16)
    float* c_w1p_score = nullptr;                    // the same image kept fresh for the scoring mode (critic_packed_stale)
    float* c_w1p = nullptr;                          // critic first conv packed [Co][Kpad] (k = tap * Ci + channel), tcg_pack_smallci_weights
    float* c_wT = nullptr; bool c_wT_stale = true;   // critic conv kernels D2..D4 with (Ci, Co) swapped, at the offsets of c_params
    float* g_wfoldT[3] = {};                         // folded generator kernels transposed: [8 phases][8 taps][Co][Ci]
    float* g_denseT = nullptr;                       // Dense kernel transposed [N][K]
    float* g_w4p = nullptr;                          // output conv: [32 taps (27 + zeros)][64] followed by its transpose [64][32]
    bool g_tcw_stale = true;
    // data-parallel exchange over NVLink peer memory (dp_peer.cu): the peers' gradient buffers / flag arrays mapped through CUDA IPC
    struct RdgPeer {
        int world = 0, rank = 0;
        void* mapped[8][3] = {};                     // what cudaIpcOpenMemHandle returned, per rank: generator grads, critic grads, flags
        float** d_grads[2] = {};                     // device arrays [world] of gradient buffer pointers (own buffer at index rank)
        unsigned** d_flags = nullptr;                // device array [world] of flag-array pointers
        unsigned* flags = nullptr;                   // this rank's flag array (exported)
        unsigned* epoch = nullptr;                   // barrier count, device-resident (graph replay)
    } peer;
    TcgArena splitk_arena[2];                        // split-K partial slices: [0] critic passes + generator backward (one stream), [1] generator-step forward
    RdgTrainState* tstate = nullptr;                 // device-resident Philox / Adam step counters (replayable CUDA graphs)
    float* rnd_buf = nullptr; size_t rnd_cap = 0;    // latent / alpha / dropout masks drawn on the device (rdg_*_step_dev)
    // forward workspace
    void* ws = nullptr; size_t ws_bytes = 0; size_t per_sample16 = 0, per_sample32 = 0;
    int* flag_dev = nullptr;
    // host-buffer pipeline
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
    // training: filter gradients run on a side stream next to the backward-data chain (train.cu)
    // [0] filter gradients of the main chain, [1] the gradient-penalty chain of the critic step, [2] its filter gradients
    cudaStream_t s_aux[3] = {}; cudaEvent_t ev_fork[3] = {}, ev_join[3] = {};
    cudaEvent_t ev_in[2] = {}, ev_comp[2] = {}, ev_out[2] = {};
    // pageable result buffers: ring of pinned staging slots drained by host callbacks on a fourth stream (rdg_generate_host)
    float* stage_out[3] = {}; cudaStream_t s_host = nullptr; cudaEvent_t ev_stage[3] = {}, ev_host[3] = {};
    float* small_pin = nullptr; float* small_dev = nullptr;   // single-stream path of rdg_generate_host for calls of <= 64 scenarios
    float* e2e_lat[2] = {}; float* e2e_out[2] = {}; float* e2e_cond = nullptr; size_t e2e_cond_cap = 0;
    // rdg_generate_stats_host staging (grow-only): observations, area means, CRPS area means, per-chunk CRPS field
    float* st_buf[4] = {}; size_t st_cap[4] = {};
    // instrumentation: kernel launch counter and optional per-layer CUDA-event timing
    long long launches = 0;
    bool prof_on = false;
    struct ProfRec { cudaEvent_t a, b; int layer; int units; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
};

using RdgPeer = rdg_ctx::RdgPeer;
void rdg_peer_destroy(rdg_ctx* c);
ConvGeom rdg_gen_conv_geom(const rdg_ctx* c, int layer, int B);
ConvGeom rdg_gen_dense_geom(const rdg_ctx* c, int B);
ConvGeom rdg_critic_conv_geom(const rdg_ctx* c, int layer, int B);
ConvGeom rdg_critic_dense_geom(const rdg_ctx* c, int B);
int rdg_repack_generator(rdg_ctx* c, cudaStream_t st, int kinds);   // kinds: bit 0 = bf16, bit 1 = fp16
static inline int rdg_kind_bit(int mode) { return mode == 1 /* RDG_MODE_BF16 */ ? 1 : 2; }
int rdg_refold32(rdg_ctx* c, cudaStream_t st);
int rdg_refresh_train_weights(rdg_ctx* c, int which, cudaStream_t st);   // train.cu: derived images of train mode 1 (which: 0 gen, 1 critic)      // refresh g_wfold32 if the generator weights changed
