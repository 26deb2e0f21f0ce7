// C ABI (include/rdg_b200.h): context, weight packing, generator / critic forward orchestration
// and the host-buffer end-to-end pipeline.  Training steps live in train.cu.
#include "rdg_common.cuh"
#include "gen_tc.h"
#include "ctx.h"
#include "tcg.h"
#include "../../include/rdg_b200.h"

#include <cstdarg>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cstdlib>

// ------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
void rdg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* rdg_last_error(void) { return g_err; }
extern "C" int rdg_version(void) { return 100; }

// ------------------------------------------------------------------ geometry helpers
static void tf_same(int i, int k, int s, int* o, int* before) {
    *o = (i + s - 1) / s;
    int p = std::max((*o - 1) * s + k - i, 0);
    *before = p / 2;
}

ConvGeom rdg_gen_conv_geom(const rdg_ctx* c, int layer, int B) {   // layer 0..2 upsampled convs, 3 = output conv
    const int s = c->nd / 8;
    static const int cin[4] = {256, 256, 128, 64}, cout[4] = {256, 128, 64, 1};
    ConvGeom g{};
    g.B = B;
    const int f = 1 << layer;
    if (layer < 3) { g.Ti = 3 * f; g.Hi = s * f; g.Wi = s * f; g.up = 1; }
    else { g.Ti = 24; g.Hi = c->nd; g.Wi = c->nd; g.up = 0; }
    g.Ci = cin[layer]; g.Co = cout[layer];
    g.To = layer < 3 ? g.Ti * 2 : g.Ti; g.Ho = layer < 3 ? g.Hi * 2 : g.Hi; g.Wo = layer < 3 ? g.Wi * 2 : g.Wi;
    g.KT = g.KH = g.KW = 3; g.stride = 1; g.pt = g.ph = g.pw = 1;
    return g;
}

ConvGeom rdg_gen_dense_geom(const rdg_ctx* c, int B) {
    ConvGeom g{};
    g.B = B; g.Ti = g.Hi = g.Wi = 1; g.To = g.Ho = g.Wo = 1;
    g.Ci = RDG_LATENT + c->nd * c->nd * c->ncond; g.Co = 256 * (c->nd / 8) * (c->nd / 8) * 3;
    g.KT = g.KH = g.KW = 1; g.stride = 1;
    return g;
}

ConvGeom rdg_critic_conv_geom(const rdg_ctx* c, int layer, int B) {   // layer 0..3
    static const int cout[4] = {64, 128, 256, 256};
    int dims[3] = {RDG_NHOURS, c->nd, c->nd};
    int cin = 1 + c->ncond;
    ConvGeom g{};
    for (int l = 0; l <= layer; ++l) {
        g = ConvGeom{};
        g.B = B; g.Ti = dims[0]; g.Hi = dims[1]; g.Wi = dims[2]; g.Ci = cin; g.Co = cout[l];
        g.KT = g.KH = g.KW = 3; g.stride = 2;
        if (l == 0) {   // padding='valid'  gan_train_cwgangp_pixelnorm.py:286-287
            g.To = (dims[0] - 3) / 2 + 1; g.Ho = (dims[1] - 3) / 2 + 1; g.Wo = (dims[2] - 3) / 2 + 1;
        } else {        // TF 'same', asymmetric for even inputs
            tf_same(dims[0], 3, 2, &g.To, &g.pt); tf_same(dims[1], 3, 2, &g.Ho, &g.ph); tf_same(dims[2], 3, 2, &g.Wo, &g.pw);
        }
        dims[0] = g.To; dims[1] = g.Ho; dims[2] = g.Wo; cin = g.Co;
    }
    return g;
}

ConvGeom rdg_critic_dense_geom(const rdg_ctx* c, int B) {
    ConvGeom l3 = rdg_critic_conv_geom(c, 3, B);
    ConvGeom g{};
    g.B = B; g.Ti = g.Hi = g.Wi = 1; g.To = g.Ho = g.Wo = 1;
    g.Ci = l3.To * l3.Ho * l3.Wo * 256; g.Co = 1; g.KT = g.KH = g.KW = 1; g.stride = 1;
    return g;
}

static void gen_shapes(const rdg_ctx* c, size_t* sz) {
    ConvGeom d = rdg_gen_dense_geom(c, 1);
    sz[0] = (size_t)d.Ci * d.Co; sz[1] = d.Co;
    static const int cin[4] = {256, 256, 128, 64}, cout[4] = {256, 128, 64, 1};
    for (int l = 0; l < 4; ++l) { sz[2 + 2 * l] = (size_t)27 * cin[l] * cout[l]; sz[3 + 2 * l] = cout[l]; }
}
static void critic_shapes(const rdg_ctx* c, size_t* sz) {
    for (int l = 0; l < 4; ++l) {
        ConvGeom g = rdg_critic_conv_geom(c, l, 1);
        sz[2 * l] = (size_t)27 * g.Ci * g.Co; sz[2 * l + 1] = g.Co;
    }
    ConvGeom d = rdg_critic_dense_geom(c, 1);
    sz[8] = d.Ci; sz[9] = 1;
}

// per-sample element counts of the generator activations
static size_t gen_act_elems(const rdg_ctx* c, int layer) {   // 0 = dense out, 1..3 conv outs
    const int s = c->nd / 8;
    static const int ch[4] = {256, 256, 128, 64};
    const int f = 1 << layer;
    return (size_t)3 * f * s * f * s * f * ch[layer];
}

// ------------------------------------------------------------------ context
extern "C" int rdg_ctx_create(rdg_ctx** out, int device, int nd, int ncond, int max_chunk) {
    if (!out || nd < 16 || (nd & (nd - 1)) || nd > 256 || ncond < 1) { rdg_set_error("rdg_ctx_create: bad arguments"); return RDG_E_BADARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) { rdg_set_error("no CUDA device %d", device); return RDG_E_NODEVICE; }
    RDG_CUDA(cudaSetDevice(device));
    rdg_ctx* c = new rdg_ctx();
    c->device = device; c->nd = nd; c->ncond = ncond;
    cudaDeviceProp prop;
    RDG_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    {   // the split-K partial buffers come from the stream-ordered pool: keep freed blocks cached instead of returning them
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    if (prop.major != 10) { rdg_set_error("device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor); delete c; return RDG_E_NODEVICE; }
    if (max_chunk <= 0) max_chunk = nd == 16 ? 32 * c->sm_count : std::max(32, 2 * c->sm_count);
    c->max_chunk = max_chunk;
    {
        const char* e = getenv("RDG_CONV3");
        c->conv3_planes = nd == 16 && !(e && strcmp(e, "tiles") == 0);
        c->conv3_logits = !(e && strcmp(e, "planes_p") == 0);   // planes_p: tap products through HBM + gather kernel
        const char* e2 = getenv("RDG_DENSE");
        c->dense_tc = nd == 16 && ncond == 1 && !(e2 && strcmp(e2, "simt") == 0);
        c->dense_big_tc = !c->dense_tc && !(e2 && strcmp(e2, "simt") == 0);
    }
    gen_shapes(c, c->g_size);
    critic_shapes(c, c->c_size);
    c->g_total = c->c_total = 0;
    for (int i = 0; i < 10; ++i) { c->g_off[i] = c->g_total; c->g_total += (c->g_size[i] + 3) / 4 * 4; }
    for (int i = 0; i < 10; ++i) { c->c_off[i] = c->c_total; c->c_total += (c->c_size[i] + 3) / 4 * 4; }
    RDG_CUDA(cudaMalloc(&c->g_params, c->g_total * 4));
    RDG_CUDA(cudaMalloc(&c->c_params, c->c_total * 4));
    RDG_CUDA(cudaMemset(c->g_params, 0, c->g_total * 4));
    RDG_CUDA(cudaMemset(c->c_params, 0, c->c_total * 4));
    // workspace: sized for the 16-bit path at max_chunk samples
    size_t per16 = 0, per32 = 0;
    ConvGeom d = rdg_gen_dense_geom(c, 1);
    per16 += (size_t)(d.Ci + d.Co) * 4;
    per32 += (size_t)(d.Ci + d.Co) * 4;
    for (int l = 0; l < 4; ++l) { per16 += gen_act_elems(c, l) * 2; per32 += (l ? gen_act_elems(c, l) * 4 : 0); }
    per16 += gen_act_elems(c, 3) / 64 * 32 * 4;   // f32 tap products P of the fused output conv
    per32 += gen_act_elems(c, 3) * 4 + 2048;      // phase-major staging of the folded FP32 convs (+ alignment slack)
    c->per_sample16 = per16; c->per_sample32 = per32;
    c->ws_bytes = per16 * (size_t)max_chunk + 4096 * 8;
    RDG_CUDA(cudaMalloc(&c->ws, c->ws_bytes));
    RDG_CUDA(cudaMalloc(&c->flag_dev, sizeof(int)));
    RDG_CUDA(cudaMemset(c->flag_dev, 0, sizeof(int)));
    RDG_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    RDG_CUDA(cudaStreamCreateWithFlags(&c->s_comp, cudaStreamNonBlocking));
    RDG_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        RDG_CUDA(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
        RDG_CUDA(cudaEventCreateWithFlags(&c->ev_comp[i], cudaEventDisableTiming));
        RDG_CUDA(cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming));
    }
    *out = c;
    return 0;
}

extern "C" void rdg_ctx_destroy(rdg_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    cudaFree(c->g_params); cudaFree(c->c_params); cudaFree(c->ws); cudaFree(c->flag_dev);
    cudaFree(c->g_grads); cudaFree(c->g_m); cudaFree(c->g_v);
    cudaFree(c->c_grads); cudaFree(c->c_m); cudaFree(c->c_v);
    for (int k = 0; k < 2; ++k) for (int l = 0; l < 3; ++l) cudaFree(c->g_wpack[k][l]);
    for (int i = 0; i < 4; ++i) cudaFree(c->st_buf[i]);
    for (int i = 0; i < 3; ++i) cudaFree(c->g_wfold32[i]);
    for (int k = 0; k < 2; ++k) cudaFree(c->g_wpack_dense_big[k]);
    for (int k = 0; k < 2; ++k) for (int l = 0; l < 3; ++l) cudaFree(c->c_wpack[k][l]);
    for (int k = 0; k < 2; ++k) { cudaFree(c->g_w4pack[k]); cudaFree(c->g_wpack_planes[k]); cudaFree(c->g_wpack_dense[k]); }
    for (int i = 0; i < 2; ++i) {
        cudaFree(c->e2e_lat[i]); cudaFree(c->e2e_out[i]);
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_comp[i]) cudaEventDestroy(c->ev_comp[i]);
        if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
    }
    cudaFree(c->e2e_cond);
    if (c->small_pin) cudaFreeHost(c->small_pin);
    cudaFree(c->small_dev);
    for (int i = 0; i < 3; ++i) {
        if (c->stage_out[i]) cudaFreeHost(c->stage_out[i]);
        if (c->ev_stage[i]) cudaEventDestroy(c->ev_stage[i]);
        if (c->ev_host[i]) cudaEventDestroy(c->ev_host[i]);
    }
    if (c->s_host) cudaStreamDestroy(c->s_host);
    rdg_peer_destroy(c);
    cudaFree(c->splitk_arena[0].buf); cudaFree(c->splitk_arena[1].buf);
    cudaFree(c->train_ws); cudaFree(c->train_ws_gen); cudaFree(c->rnd_buf_gen); cudaFree(c->train_ws1); cudaFree(c->rnd_buf1);
    cudaFree(c->c_wT); cudaFree(c->c_w1p); cudaFree(c->c_w1p_score); cudaFree(c->c_w1q_score); cudaFree(c->g_denseT); cudaFree(c->g_w4p); cudaFree(c->tstate); cudaFree(c->rnd_buf);
    for (int i = 0; i < 3; ++i) cudaFree(c->g_wfoldT[i]);
    for (int i = 0; i < 3; ++i) {
        if (c->s_aux[i]) cudaStreamDestroy(c->s_aux[i]);
        if (c->ev_fork[i]) cudaEventDestroy(c->ev_fork[i]);
        if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
    }
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_comp) cudaStreamDestroy(c->s_comp);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    delete c;
}

extern "C" int rdg_ctx_info(const rdg_ctx* c, int* nd, int* ncond, int* max_chunk, int* sm_count) {
    if (!c) return RDG_E_BADARG;
    if (nd) *nd = c->nd;
    if (ncond) *ncond = c->ncond;
    if (max_chunk) *max_chunk = c->max_chunk;
    if (sm_count) *sm_count = c->sm_count;
    return 0;
}
extern "C" size_t rdg_ctx_workspace_bytes(const rdg_ctx* c) { return c ? c->ws_bytes : 0; }

// ------------------------------------------------------------------ weight packing
static int upload_params(float* dev, const size_t* off, const size_t* size, const float* const* tensors,
                         const size_t* sizes, int n, const char* what) {
    if (n != 10) { rdg_set_error("%s: expected 10 tensors, got %d", what, n); return RDG_E_BADARG; }
    for (int i = 0; i < 10; ++i)
        if (sizes[i] != size[i]) { rdg_set_error("%s: tensor %d has %zu elements, expected %zu", what, i, sizes[i], size[i]); return RDG_E_BADARG; }
    for (int i = 0; i < 10; ++i) RDG_CUDA(cudaMemcpy(dev + off[i], tensors[i], size[i] * 4, cudaMemcpyHostToDevice));
    return 0;
}

// FP32 folded kernels for the SIMT path (FP32 inference mode, training step)
int rdg_refold32(rdg_ctx* c, cudaStream_t st) {
    static const int cin[3] = {256, 256, 128}, cout[3] = {256, 128, 64};
    if (!c->fold32_stale) return 0;
    for (int l = 0; l < 3; ++l) {
        if (!c->g_wfold32[l]) RDG_CUDA(cudaMalloc(&c->g_wfold32[l], folded_weight_elems(cin[l], cout[l]) * sizeof(float)));
        int r = folded_pack_f32(c->g_params + c->g_off[2 + 2 * l], c->g_wfold32[l], cin[l], cout[l], st);
        if (r) return r;
    }
    c->launches += 3;
    c->fold32_stale = false;
    return 0;
}

// (Re)build the folded + swizzled 16-bit operand tiles from the f32 master weights, on the device.
int rdg_repack_generator(rdg_ctx* c, cudaStream_t st, int kinds) {
    static const int cin[3] = {256, 256, 128}, cout[3] = {256, 128, 64};
    for (int k = 0; k < 2; ++k) {
        if (!(kinds & (1 << k))) continue;
        const int hk = k == 0 ? RDG_HALF_BF16 : RDG_HALF_FP16;
        for (int l = 0; l < 3; ++l) {
            if (!c->g_wpack[k][l]) RDG_CUDA(cudaMalloc(&c->g_wpack[k][l], (size_t)64 * cin[l] * cout[l] * 2));
            int r = pack_folded_weights(hk, c->g_params + c->g_off[2 + 2 * l], c->g_wpack[k][l], cin[l], cout[l], st);
            if (r) return r;
        }
        if (c->conv3_planes) {
            if (!c->g_wpack_planes[k]) RDG_CUDA(cudaMalloc(&c->g_wpack_planes[k], (size_t)64 * 128 * 64 * 2));
            int r2 = pack_folded_weights_planes(hk, c->g_params + c->g_off[6], c->g_wpack_planes[k], st);
            if (r2) return r2;
            c->launches += 1;
        }
        if (c->dense_big_tc) {
            const ConvGeom dg = rdg_gen_dense_geom(c, 1);
            if (!c->g_wpack_dense_big[k]) RDG_CUDA(cudaMalloc(&c->g_wpack_dense_big[k], tc_dense_big_pack_bytes(dg.Ci, dg.Co)));
            int r2 = pack_dense_big_weights(hk, c->g_params + c->g_off[0], c->g_wpack_dense_big[k], dg.Ci, dg.Co, st);
            if (r2) return r2;
            c->launches += 1;
        }
        if (c->dense_tc) {
            if (!c->g_wpack_dense[k]) RDG_CUDA(cudaMalloc(&c->g_wpack_dense[k], tc_dense_pack_bytes()));
            int r2 = pack_dense_weights(hk, c->g_params + c->g_off[0], c->g_wpack_dense[k], st);
            if (r2) return r2;
            c->launches += 1;
        }
        if (!c->g_w4pack[k]) RDG_CUDA(cudaMalloc(&c->g_w4pack[k], RDG_W4PACK_BYTES));
        int r = pack_w4_tile(hk, c->g_params + c->g_off[8], c->g_w4pack[k], st);
        if (r) return r;
    }
    c->launches += 10;
    c->gen_stale_kinds &= ~kinds;
    return 0;
}

extern "C" int rdg_generator_set_weights(rdg_ctx* c, const float* const* tensors, const size_t* sizes, int n) {
    if (!c || !tensors || !sizes) return RDG_E_BADARG;
    RDG_CUDA(cudaSetDevice(c->device));
    int r = upload_params(c->g_params, c->g_off, c->g_size, tensors, sizes, n, "generator weights");
    if (r) return r;
    c->fold32_stale = true; c->g_tcw_stale = true; c->gen_stale_kinds = 3;
    r = rdg_repack_generator(c, nullptr, 3);
    if (r) return r;
    c->gen_ready = true;
    if (c->train_mode == 1 && (r = rdg_refresh_train_weights(c, 0, nullptr))) return r;
    RDG_CUDA(cudaDeviceSynchronize());
    return 0;
}

static int download_params(const float* dev, const size_t* off, const size_t* size, float* const* tensors,
                           const size_t* sizes, int n, const char* what) {
    if (n != 10) { rdg_set_error("%s: expected 10 tensors", what); return RDG_E_BADARG; }
    for (int i = 0; i < 10; ++i)
        if (sizes[i] != size[i]) { rdg_set_error("%s: tensor %d size mismatch", what, i); return RDG_E_BADARG; }
    RDG_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < 10; ++i) RDG_CUDA(cudaMemcpy(tensors[i], dev + off[i], size[i] * 4, cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int rdg_generator_get_weights(rdg_ctx* c, float* const* tensors, const size_t* sizes, int n) {
    if (!c) return RDG_E_BADARG;
    RDG_CUDA(cudaSetDevice(c->device));
    return download_params(c->g_params, c->g_off, c->g_size, tensors, sizes, n, "generator weights");
}
extern "C" int rdg_critic_set_weights(rdg_ctx* c, const float* const* tensors, const size_t* sizes, int n) {
    if (!c || !tensors || !sizes) return RDG_E_BADARG;
    RDG_CUDA(cudaSetDevice(c->device));
    int r = upload_params(c->c_params, c->c_off, c->c_size, tensors, sizes, n, "critic weights");
    if (r) return r;
    c->critic_ready = true;
    c->critic_packed_stale = true; c->c_wT_stale = true;
    if (c->train_mode == 1 && (r = rdg_refresh_train_weights(c, 1, nullptr))) return r;
    RDG_CUDA(cudaDeviceSynchronize());
    return 0;
}
extern "C" int rdg_critic_get_weights(rdg_ctx* c, float* const* tensors, const size_t* sizes, int n) {
    if (!c) return RDG_E_BADARG;
    RDG_CUDA(cudaSetDevice(c->device));
    return download_params(c->c_params, c->c_off, c->c_size, tensors, sizes, n, "critic weights");
}

// ------------------------------------------------------------------ instrumentation
namespace {
struct ProfScope {
    rdg_ctx* c; cudaStream_t st; rdg_ctx::ProfRec rec; bool on;
    ProfScope(rdg_ctx* c_, cudaStream_t st_, int layer, int units, int nlaunch) : c(c_), st(st_), on(c_->prof_on) {
        c->launches += nlaunch;
        if (!on) return;
        auto get = [&]() { cudaEvent_t e; if (c->prof_pool.empty()) cudaEventCreate(&e); else { e = c->prof_pool.back(); c->prof_pool.pop_back(); } return e; };
        rec.a = get(); rec.b = get(); rec.layer = layer; rec.units = units;
        cudaEventRecord(rec.a, st);
    }
    ~ProfScope() { if (on) { cudaEventRecord(rec.b, st); c->prof_recs.push_back(rec); } }
};
}  // namespace

extern "C" int rdg_profile_enable(rdg_ctx* c, int on) {
    if (!c) return RDG_E_BADARG;
    c->prof_on = on != 0;
    return 0;
}
extern "C" long long rdg_launch_count(const rdg_ctx* c) { return c ? c->launches : -1; }
// Sums the CUDA-event durations recorded since the last collect, per layer id
// (0 input concat, 1 dense, 2 f32->16bit, 3..5 upsampled convs, 6 output conv + softmax, 7 pixelnorm (fp32 mode)).
extern "C" int rdg_profile_collect(rdg_ctx* c, double* ms_sum8, long long* launches8, long long* units8) {
    if (!c) return RDG_E_BADARG;
    RDG_CUDA(cudaSetDevice(c->device));
    for (int i = 0; i < 8; ++i) { if (ms_sum8) ms_sum8[i] = 0; if (launches8) launches8[i] = 0; if (units8) units8[i] = 0; }
    for (auto& r : c->prof_recs) {
        RDG_CUDA(cudaEventSynchronize(r.b));
        float ms = 0.f;
        RDG_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
        if (r.layer >= 0 && r.layer < 8) {
            if (ms_sum8) ms_sum8[r.layer] += ms;
            if (launches8) launches8[r.layer] += 1;
            if (units8) units8[r.layer] += r.units;
        }
        c->prof_pool.push_back(r.a); c->prof_pool.push_back(r.b);
    }
    c->prof_recs.clear();
    return 0;
}

// ------------------------------------------------------------------ generator forward
// One chunk of n samples starting at global sample index b_off (for the cond lookup).
static int gen_forward_chunk(rdg_ctx* c, const float* latent, const float* cond, int spc, int b_off, float* out, int n,
                             int mode, int out_kind, float norm_scale, int* flag, cudaStream_t st) {
    const int nd = c->nd, s = nd / 8;
    const ConvGeom dg = rdg_gen_dense_geom(c, n);
    uint8_t* p = reinterpret_cast<uint8_t*>(c->ws);
    auto take = [&](size_t bytes) { void* r = p; p += (bytes + 255) / 256 * 256; return r; };
    int r;
    float* x0 = nullptr;
    float* d0 = nullptr;
    const bool front_tc = mode != RDG_MODE_FP32 && c->dense_tc;
    const bool front_big = mode != RDG_MODE_FP32 && c->dense_big_tc;
    if (!front_tc && !front_big) {
        x0 = (float*)take((size_t)n * dg.Ci * 4);
        d0 = (float*)take((size_t)n * dg.Co * 4);
        { ProfScope ps(c, st, 0, n, 1);
          if ((r = ew_assemble_gen_input(latent, cond, spc, b_off, x0, n, nd * nd * c->ncond, st))) return r; }
        { ProfScope ps(c, st, 1, n, 1);
          if ((r = simt_conv_fwd(x0, c->g_params + c->g_off[0], c->g_params + c->g_off[1], d0, dg, ACT_LRELU, nullptr, 1.f, st))) return r; }
    }
    const float* w4 = c->g_params + c->g_off[8];
    const float* b4 = c->g_params + c->g_off[9];
    if (mode == RDG_MODE_FP32) {
        const float* cur = d0;
        if ((r = rdg_refold32(c, st))) return r;
        float* scratch = (float*)take((size_t)n * gen_act_elems(c, 3) * 4);      // phase-major staging of the largest layer
        for (int l = 0; l < 3; ++l) {
            ConvGeom g = rdg_gen_conv_geom(c, l, n);
            float* y = (float*)take((size_t)n * gen_act_elems(c, l + 1) * 4);
            { ProfScope ps(c, st, 3 + l, n, 1);     // upsample-folded: 8 phase convs with 2^3 taps on the low-res grid
              if ((r = folded_conv_fwd(cur, c->g_wfold32[l], c->g_params + c->g_off[3 + 2 * l], y, scratch, g, st))) return r; }
            { ProfScope ps(c, st, 7, n, 1);
              if ((r = ew_pixelnorm(y, y, (long long)n * g.To * g.Ho * g.Wo, g.Co, 1, st))) return r; }
            cur = y;
        }
        ProfScope ps(c, st, 6, n, 1);
        return conv_out_softmax(RDG_HALF_F32, cur, w4, b4, out, cond, n, nd, spc, b_off, c->ncond, norm_scale,
                                out_kind == RDG_OUT_MM, flag, st);
    }
    const int hk = mode == RDG_MODE_BF16 ? RDG_HALF_BF16 : RDG_HALF_FP16;
    void* h = take((size_t)n * gen_act_elems(c, 0) * 2);
    if (front_tc) {
        ProfScope ps(c, st, 1, n, 1);
        if ((r = tc_dense_lrelu(hk, latent, cond, spc, b_off, c->g_wpack_dense[hk == RDG_HALF_BF16 ? 0 : 1], c->g_params + c->g_off[1], h, n, st))) return r;
    } else if (front_big) {
        ProfScope ps(c, st, 1, n, 2);
        void* x16 = take(tc_dense_big_input_bytes(n, dg.Ci));
        if ((r = tc_dense_big_lrelu(hk, latent, cond, spc, b_off, c->g_wpack_dense_big[hk == RDG_HALF_BF16 ? 0 : 1], c->g_params + c->g_off[1],
                                    x16, h, n, dg.Ci, dg.Co, st))) return r;
    } else {
        ProfScope ps(c, st, 2, n, 1);
        if ((r = f32_to_half(hk, d0, h, (long long)n * dg.Co, st))) return r;
    }
    static const int cin[3] = {256, 256, 128}, cout[3] = {256, 128, 64};
    const int wk = hk == RDG_HALF_BF16 ? 0 : 1;
    float* pbuf = nullptr;
    // resident-plane kernel: the output conv is summed on chip, `out` carries the fixed-point logits to the softmax
    const bool logits_on_chip = c->conv3_planes && c->conv3_logits;
    for (int l = 0; l < 3; ++l) {
        const int f = 1 << l;
        void* y = nullptr;
        if (l < 2) y = take((size_t)n * gen_act_elems(c, l + 1) * 2);
        else if (!logits_on_chip) pbuf = (float*)take((size_t)n * gen_act_elems(c, 3) / 64 * 32 * 4);
        { ProfScope ps(c, st, 3 + l, n, 1);
          if (l == 2 && c->conv3_planes)
              r = tc_upconv64_planes(hk, h, c->g_wpack_planes[wk], c->g_params + c->g_off[7], nullptr, c->g_w4pack[wk], pbuf,
                                     logits_on_chip ? reinterpret_cast<int*>(out) : nullptr, flag, n, 3 * f, c->sm_count, st);
          else
              r = tc_upconv_pixelnorm(hk, h, c->g_wpack[wk][l], c->g_params + c->g_off[3 + 2 * l], y,
                                      l == 2 ? c->g_w4pack[wk] : nullptr, l == 2 ? pbuf : nullptr, n,
                                      3 * f, s * f, s * f, cin[l], cout[l], c->sm_count, st);
          if (r) return r; }
        h = y;
    }
    ProfScope ps(c, st, 6, n, 1);
    if (logits_on_chip)
        return softmax_fixed_inplace(out, c->g_w4pack[wk], b4, cond, n, spc, b_off, c->ncond, norm_scale, out_kind == RDG_OUT_MM, flag, st);
    return gather_softmax(pbuf, b4, out, cond, n, nd, spc, b_off, c->ncond, norm_scale, out_kind == RDG_OUT_MM, flag, st);
}

static int chunk_for_mode(const rdg_ctx* c, int mode) {
    if (mode != RDG_MODE_FP32) return c->max_chunk;
    size_t n = (c->ws_bytes - 4096 * 8) / c->per_sample32;
    return (int)std::max<size_t>(1, std::min<size_t>(n, (size_t)c->max_chunk));
}

extern "C" int rdg_generator_forward(rdg_ctx* c, const float* latent_dev, const float* cond_dev, int spc, float* out_dev,
                                     int B, int mode, int out_kind, float norm_scale, int* flag_dev, void* stream) {
    if (!c || B < 0 || spc < 1 || mode < 0 || mode > 2) { rdg_set_error("rdg_generator_forward: bad arguments"); return RDG_E_BADARG; }
    if (!c->gen_ready) { rdg_set_error("generator weights not set"); return RDG_E_NOWEIGHT; }
    RDG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (mode != RDG_MODE_FP32 && (c->gen_stale_kinds & rdg_kind_bit(mode))) { int r = rdg_repack_generator(c, st, rdg_kind_bit(mode)); if (r) return r; }
    const int chunk = chunk_for_mode(c, mode);
    const size_t plane = (size_t)RDG_NHOURS * c->nd * c->nd;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        int n = std::min(chunk, B - b0);
        int r = gen_forward_chunk(c, latent_dev + (size_t)b0 * RDG_LATENT, cond_dev, spc, b0, out_dev + (size_t)b0 * plane, n,
                                  mode, out_kind, norm_scale, flag_dev, st);
        if (r) return r;
    }
    return 0;
}

namespace {
struct HostCopyJob { void* dst; const void* src; size_t bytes; };
void CUDART_CB host_copy_cb(void* p) {
    HostCopyJob* j = static_cast<HostCopyJob*>(p);
    memcpy(j->dst, j->src, j->bytes);
    delete j;
}
bool is_pinned_host(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}
void drain_host_pipeline(rdg_ctx* c) {
    cudaStreamSynchronize(c->s_h2d); cudaStreamSynchronize(c->s_comp); cudaStreamSynchronize(c->s_d2h);
    if (c->s_host) cudaStreamSynchronize(c->s_host);
}
}  // namespace

// Stream contract of the *_host calls: they run on the context's own three (four) streams and block until the result is in the
// caller's buffer.  Weight-mutating calls (set_weights, rdg_adam_apply*) run on the CALLER's stream, so a *_host call that finds
// the packed generator weights stale first waits for the whole device (the only ordering available without the caller's stream).
extern "C" int rdg_generate_host(rdg_ctx* c, const float* latent_host, const float* cond_host, int spc, float* out_host,
                                 long long B, int mode, int out_kind, float norm_scale) {
    if (!c || B < 0 || spc < 1 || mode < 0 || mode > 2) { rdg_set_error("rdg_generate_host: bad arguments"); return RDG_E_BADARG; }
    if (!c->gen_ready) { rdg_set_error("generator weights not set"); return RDG_E_NOWEIGHT; }
    if (B == 0) return 0;
    RDG_CUDA(cudaSetDevice(c->device));
    if (c->gen_stale_kinds || c->fold32_stale) RDG_CUDA(cudaDeviceSynchronize());
    if (mode != RDG_MODE_FP32 && (c->gen_stale_kinds & rdg_kind_bit(mode))) { int r = rdg_repack_generator(c, c->s_comp, rdg_kind_bit(mode)); if (r) return r; }
    const int chunk = chunk_for_mode(c, mode);
    const size_t plane = (size_t)RDG_NHOURS * c->nd * c->nd;
    const size_t ncf = (size_t)c->nd * c->nd * c->ncond;
    const long long n_cond = (B + spc - 1) / spc;
    // Small calls (example.py: 10 scenarios): one stream, inputs and result through small pinned buffers owned by the context
    // (a cudaMemcpyAsync from / to pageable memory is staged synchronously by the driver, ~10 us each, and the three-stream
    // hand-offs of the pipeline below cost more than they hide for a single chunk).
    constexpr long long kSmall = 64;
    if (B <= kSmall && B <= chunk) {
        const size_t in_f = (size_t)kSmall * RDG_LATENT + (size_t)kSmall * ncf, out_f = (size_t)kSmall * plane;
        if (!c->small_pin) {
            RDG_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&c->small_pin), (in_f + out_f + 16) * 4, cudaHostAllocDefault));
            RDG_CUDA(cudaMalloc(&c->small_dev, (in_f + out_f) * 4));
        }
        float* h_lat = c->small_pin; float* h_cond = h_lat + kSmall * RDG_LATENT; float* h_out = c->small_pin + in_f;
        float* d_lat = c->small_dev; float* d_cond = d_lat + kSmall * RDG_LATENT; float* d_out = c->small_dev + in_f;
        memcpy(h_lat, latent_host, (size_t)B * RDG_LATENT * 4);
        memcpy(h_cond, cond_host, (size_t)n_cond * ncf * 4);
        RDG_CUDA(cudaMemsetAsync(c->flag_dev, 0, sizeof(int), c->s_comp));
        RDG_CUDA(cudaMemcpyAsync(d_lat, h_lat, ((size_t)kSmall * RDG_LATENT + (size_t)n_cond * ncf) * 4, cudaMemcpyHostToDevice, c->s_comp));
        int r = gen_forward_chunk(c, d_lat, d_cond, spc, 0, d_out, (int)B, mode, out_kind, norm_scale, c->flag_dev, c->s_comp);
        if (r) { cudaStreamSynchronize(c->s_comp); return r; }
        RDG_CUDA(cudaMemcpyAsync(h_out, d_out, (size_t)B * plane * 4, cudaMemcpyDeviceToHost, c->s_comp));
        int* h_flag = reinterpret_cast<int*>(h_out + (size_t)B * plane);
        RDG_CUDA(cudaMemcpyAsync(h_flag, c->flag_dev, sizeof(int), cudaMemcpyDeviceToHost, c->s_comp));
        RDG_CUDA(cudaStreamSynchronize(c->s_comp));
        memcpy(out_host, h_out, (size_t)B * plane * 4);
        if (*h_flag) { rdg_set_error("found nan in output of per_gridpoint_softmax"); return RDG_E_NONFINITE; }
        return 0;
    }
    for (int i = 0; i < 2; ++i) {
        if (!c->e2e_lat[i]) RDG_CUDA(cudaMalloc(&c->e2e_lat[i], (size_t)c->max_chunk * RDG_LATENT * 4));
        if (!c->e2e_out[i]) RDG_CUDA(cudaMalloc(&c->e2e_out[i], (size_t)c->max_chunk * plane * 4));
    }
    if (c->e2e_cond_cap < (size_t)n_cond * ncf) {
        cudaFree(c->e2e_cond);
        c->e2e_cond = nullptr;
        RDG_CUDA(cudaMalloc(&c->e2e_cond, (size_t)n_cond * ncf * 4));
        c->e2e_cond_cap = (size_t)n_cond * ncf;
    }
    // Pageable result buffer (e.g. the numpy array of gen.predict): a cudaMemcpyAsync into it would be staged by the driver
    // synchronously, chunk by chunk, with no overlap.  Instead the D2H copies land in a ring of three pinned staging slots and a
    // host callback on a fourth stream moves each slot into the caller's buffer while the next chunk is already on the wire.
    const bool staged = B > chunk && !is_pinned_host(out_host);
    if (staged) {
        if (!c->s_host) RDG_CUDA(cudaStreamCreateWithFlags(&c->s_host, cudaStreamNonBlocking));
        for (int i = 0; i < 3; ++i) {
            if (!c->stage_out[i]) RDG_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&c->stage_out[i]), (size_t)c->max_chunk * plane * 4, cudaHostAllocDefault));
            if (!c->ev_stage[i]) { RDG_CUDA(cudaEventCreateWithFlags(&c->ev_stage[i], cudaEventDisableTiming)); RDG_CUDA(cudaEventCreateWithFlags(&c->ev_host[i], cudaEventDisableTiming)); }
        }
    }
    RDG_CUDA(cudaMemsetAsync(c->flag_dev, 0, sizeof(int), c->s_comp));
    RDG_CUDA(cudaMemcpyAsync(c->e2e_cond, cond_host, (size_t)n_cond * ncf * 4, cudaMemcpyHostToDevice, c->s_comp));
    long long it = 0;
    for (long long b0 = 0; b0 < B; b0 += chunk, ++it) {
        const int n = (int)std::min<long long>(chunk, B - b0);
        const int k = (int)(it & 1);
        // the latent buffer k is free once the compute of chunk it-2 finished
        if (it >= 2) RDG_CUDA(cudaStreamWaitEvent(c->s_h2d, c->ev_comp[k], 0));
        RDG_CUDA(cudaMemcpyAsync(c->e2e_lat[k], latent_host + b0 * RDG_LATENT, (size_t)n * RDG_LATENT * 4, cudaMemcpyHostToDevice, c->s_h2d));
        RDG_CUDA(cudaEventRecord(c->ev_in[k], c->s_h2d));
        RDG_CUDA(cudaStreamWaitEvent(c->s_comp, c->ev_in[k], 0));
        // the output buffer k is free once the D2H of chunk it-2 finished
        if (it >= 2) RDG_CUDA(cudaStreamWaitEvent(c->s_comp, c->ev_out[k], 0));
        int r = gen_forward_chunk(c, c->e2e_lat[k], c->e2e_cond, spc, (int)b0, c->e2e_out[k], n, mode, out_kind, norm_scale,
                                  c->flag_dev, c->s_comp);
        if (r) { drain_host_pipeline(c); return r; }
        RDG_CUDA(cudaEventRecord(c->ev_comp[k], c->s_comp));
        RDG_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_comp[k], 0));
        if (!staged) {
            RDG_CUDA(cudaMemcpyAsync(out_host + b0 * plane, c->e2e_out[k], (size_t)n * plane * 4, cudaMemcpyDeviceToHost, c->s_d2h));
            RDG_CUDA(cudaEventRecord(c->ev_out[k], c->s_d2h));
        } else {
            const int slot = (int)(it % 3);
            if (it >= 3) RDG_CUDA(cudaStreamWaitEvent(c->s_d2h, c->ev_host[slot], 0));      // slot drained by its host copy
            RDG_CUDA(cudaMemcpyAsync(c->stage_out[slot], c->e2e_out[k], (size_t)n * plane * 4, cudaMemcpyDeviceToHost, c->s_d2h));
            RDG_CUDA(cudaEventRecord(c->ev_out[k], c->s_d2h));
            RDG_CUDA(cudaEventRecord(c->ev_stage[slot], c->s_d2h));
            RDG_CUDA(cudaStreamWaitEvent(c->s_host, c->ev_stage[slot], 0));
            HostCopyJob* job = new HostCopyJob{out_host + b0 * plane, c->stage_out[slot], (size_t)n * plane * 4};
            cudaError_t e = cudaLaunchHostFunc(c->s_host, host_copy_cb, job);
            if (e != cudaSuccess) { delete job; drain_host_pipeline(c); rdg_set_error("cudaLaunchHostFunc -> %s", cudaGetErrorString(e)); return (int)e; }
            RDG_CUDA(cudaEventRecord(c->ev_host[slot], c->s_host));
        }
    }
    int flag = 0;
    RDG_CUDA(cudaStreamSynchronize(c->s_d2h));
    if (staged) RDG_CUDA(cudaStreamSynchronize(c->s_host));
    RDG_CUDA(cudaMemcpyAsync(&flag, c->flag_dev, sizeof(int), cudaMemcpyDeviceToHost, c->s_comp));
    RDG_CUDA(cudaStreamSynchronize(c->s_comp));
    RDG_CUDA(cudaStreamSynchronize(c->s_h2d));
    if (flag) { rdg_set_error("found nan in output of per_gridpoint_softmax"); return RDG_E_NONFINITE; }
    return 0;
}

int rdg_stats_chunk(rdg_ctx* c, const float* fields, int n_cond, int spc, const float* obs, float* area_mean, float* crps_scratch,
                    float* crps_amean, cudaStream_t st);

// Generation + ensemble statistics, chunked by whole conditions; only the statistics are copied back.
extern "C" int rdg_generate_stats_host(rdg_ctx* c, const float* latent_host, const float* cond_host, int spc, const float* obs_host,
                                       long long B, int mode, int out_kind, float norm_scale, float* area_mean_host,
                                       float* crps_amean_host) {
    if (!c || B < 0 || spc < 1 || mode < 0 || mode > 2 || B % spc) { rdg_set_error("rdg_generate_stats_host: bad arguments (B must be n_cond * scen_per_cond)"); return RDG_E_BADARG; }
    if (crps_amean_host && !obs_host) { rdg_set_error("rdg_generate_stats_host: CRPS needs observations"); return RDG_E_BADARG; }
    if (!c->gen_ready) { rdg_set_error("generator weights not set"); return RDG_E_NOWEIGHT; }
    if (B == 0) return 0;
    RDG_CUDA(cudaSetDevice(c->device));
    if (mode != RDG_MODE_FP32 && (c->gen_stale_kinds & rdg_kind_bit(mode))) { int r = rdg_repack_generator(c, c->s_comp, rdg_kind_bit(mode)); if (r) return r; }
    const int chunk_max = chunk_for_mode(c, mode);
    if (spc > chunk_max) { rdg_set_error("rdg_generate_stats_host: %d members per condition exceed the chunk of %d samples", spc, chunk_max); return RDG_E_BADARG; }
    const int cond_per_chunk = chunk_max / spc, chunk = cond_per_chunk * spc;
    const size_t plane = (size_t)RDG_NHOURS * c->nd * c->nd;
    const size_t ncf = (size_t)c->nd * c->nd * c->ncond;
    const long long n_cond = B / spc;
    for (int i = 0; i < 2; ++i) {
        if (!c->e2e_lat[i]) RDG_CUDA(cudaMalloc(&c->e2e_lat[i], (size_t)c->max_chunk * RDG_LATENT * 4));
        if (!c->e2e_out[i]) RDG_CUDA(cudaMalloc(&c->e2e_out[i], (size_t)c->max_chunk * plane * 4));
    }
    if (c->e2e_cond_cap < (size_t)n_cond * ncf) {
        cudaFree(c->e2e_cond);
        c->e2e_cond = nullptr;
        RDG_CUDA(cudaMalloc(&c->e2e_cond, (size_t)n_cond * ncf * 4));
        c->e2e_cond_cap = (size_t)n_cond * ncf;
    }
    float* obs_dev = nullptr; float* am_dev = nullptr; float* cr_dev = nullptr; float* crps_tmp = nullptr;
    {
        const size_t want[4] = {obs_host ? (size_t)n_cond * plane : 0, area_mean_host ? (size_t)B * RDG_NHOURS : 0,
                                crps_amean_host ? (size_t)n_cond * RDG_NHOURS : 0, crps_amean_host ? (size_t)cond_per_chunk * plane : 0};
        for (int i = 0; i < 4; ++i)
            if (c->st_cap[i] < want[i]) {
                cudaFree(c->st_buf[i]); c->st_buf[i] = nullptr; c->st_cap[i] = 0;
                RDG_CUDA(cudaMalloc(&c->st_buf[i], want[i] * 4));
                c->st_cap[i] = want[i];
            }
        if (obs_host) obs_dev = c->st_buf[0];
        if (area_mean_host) am_dev = c->st_buf[1];
        if (crps_amean_host) { cr_dev = c->st_buf[2]; crps_tmp = c->st_buf[3]; }
    }
    auto cleanup = [&]() {};
#define RDG_CUDA_C(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rdg_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); cleanup(); return (int)e_; } } while (0)
    RDG_CUDA_C(cudaMemsetAsync(c->flag_dev, 0, sizeof(int), c->s_comp));
    RDG_CUDA_C(cudaMemcpyAsync(c->e2e_cond, cond_host, (size_t)n_cond * ncf * 4, cudaMemcpyHostToDevice, c->s_comp));
    if (obs_host) RDG_CUDA_C(cudaMemcpyAsync(obs_dev, obs_host, (size_t)n_cond * plane * 4, cudaMemcpyHostToDevice, c->s_comp));
    long long it = 0;
    for (long long b0 = 0; b0 < B; b0 += chunk, ++it) {
        const int n = (int)std::min<long long>(chunk, B - b0);
        const int k = (int)(it & 1);
        if (it >= 2) RDG_CUDA_C(cudaStreamWaitEvent(c->s_h2d, c->ev_comp[k], 0));     // latent buffer k free after compute of chunk it-2
        RDG_CUDA_C(cudaMemcpyAsync(c->e2e_lat[k], latent_host + b0 * RDG_LATENT, (size_t)n * RDG_LATENT * 4, cudaMemcpyHostToDevice, c->s_h2d));
        RDG_CUDA_C(cudaEventRecord(c->ev_in[k], c->s_h2d));
        RDG_CUDA_C(cudaStreamWaitEvent(c->s_comp, c->ev_in[k], 0));
        int r = gen_forward_chunk(c, c->e2e_lat[k], c->e2e_cond, spc, (int)b0, c->e2e_out[k], n, mode, out_kind, norm_scale,
                                  c->flag_dev, c->s_comp);
        if (!r) r = rdg_stats_chunk(c, c->e2e_out[k], n / spc, spc, obs_dev ? obs_dev + (b0 / spc) * plane : nullptr,
                                    am_dev ? am_dev + b0 * RDG_NHOURS : nullptr, crps_tmp, cr_dev ? cr_dev + (b0 / spc) * RDG_NHOURS : nullptr, c->s_comp);
        if (r) { cudaStreamSynchronize(c->s_comp); cudaStreamSynchronize(c->s_h2d); cleanup(); return r; }
        RDG_CUDA_C(cudaEventRecord(c->ev_comp[k], c->s_comp));
    }
    int flag = 0;
    if (area_mean_host) RDG_CUDA_C(cudaMemcpyAsync(area_mean_host, am_dev, (size_t)B * RDG_NHOURS * 4, cudaMemcpyDeviceToHost, c->s_comp));
    if (crps_amean_host) RDG_CUDA_C(cudaMemcpyAsync(crps_amean_host, cr_dev, (size_t)n_cond * RDG_NHOURS * 4, cudaMemcpyDeviceToHost, c->s_comp));
    RDG_CUDA_C(cudaMemcpyAsync(&flag, c->flag_dev, sizeof(int), cudaMemcpyDeviceToHost, c->s_comp));
    RDG_CUDA_C(cudaStreamSynchronize(c->s_comp));
    RDG_CUDA_C(cudaStreamSynchronize(c->s_h2d));
#undef RDG_CUDA_C
    cleanup();
    if (flag) { rdg_set_error("found nan in output of per_gridpoint_softmax"); return RDG_E_NONFINITE; }
    return 0;
}

extern "C" int rdg_fill_normal(float* dst_dev, long long n, uint64_t seed, uint64_t offset, void* stream) {
    if (!dst_dev || n < 0 || (offset & 3)) { rdg_set_error("rdg_fill_normal: bad arguments (offset must be a multiple of 4)"); return RDG_E_BADARG; }
    return ew_fill_normal(dst_dev, n, seed, offset, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ critic forward
extern "C" int rdg_critic_forward(rdg_ctx* c, const float* sample_dev, const float* cond_dev, const float* const* masks_dev,
                                  float* score_dev, int B, void* stream) {
    if (!c || B < 0) return RDG_E_BADARG;
    if (!c->critic_ready) { rdg_set_error("critic weights not set"); return RDG_E_NOWEIGHT; }
    RDG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    // per-sample workspace of the critic
    size_t per = (size_t)RDG_NHOURS * c->nd * c->nd * (1 + c->ncond);
    for (int l = 0; l < 4; ++l) { ConvGeom g = rdg_critic_conv_geom(c, l, 1); per += (size_t)g.To * g.Ho * g.Wo * g.Co; }
    per = per * 4 + 4 * 256;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((c->ws_bytes - 4096) / per, 1 << 20));
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int n = std::min(chunk, B - b0);
        uint8_t* p = reinterpret_cast<uint8_t*>(c->ws);
        auto take = [&](size_t bytes) { void* r = p; p += (bytes + 255) / 256 * 256; return r; };
        float* x = (float*)take((size_t)n * RDG_NHOURS * c->nd * c->nd * (1 + c->ncond) * 4);
        int r;
        if ((r = ew_critic_input(sample_dev + (size_t)b0 * RDG_NHOURS * c->nd * c->nd, cond_dev + (size_t)b0 * c->nd * c->nd * c->ncond, x, n, c->nd, c->ncond, st))) return r;
        const float* cur = x;
        for (int l = 0; l < 4; ++l) {
            ConvGeom g = rdg_critic_conv_geom(c, l, n);
            const size_t per_l = (size_t)g.To * g.Ho * g.Wo * g.Co;
            float* y = (float*)take((size_t)n * per_l * 4);
            const float* m = masks_dev ? masks_dev[l] + (size_t)b0 * per_l : nullptr;
            if ((r = simt_conv_fwd(cur, c->c_params + c->c_off[2 * l], c->c_params + c->c_off[2 * l + 1], y, g, ACT_LRELU, m, 1.f / 0.75f, st))) return r;
            cur = y;
        }
        ConvGeom d = rdg_critic_dense_geom(c, n);
        if ((r = simt_conv_fwd(cur, c->c_params + c->c_off[8], c->c_params + c->c_off[9], score_dev + b0, d, ACT_NONE, nullptr, 1.f, st))) return r;
    }
    return 0;
}

// critic([sample, cond]) in the 16-bit tensor-core scoring mode (inference: Dropout is identity)
extern "C" int rdg_critic_forward_tc(rdg_ctx* c, const float* sample_dev, const float* cond_dev, float* score_dev, int B, int mode,
                                     void* stream) {
    if (!c || B < 0 || (mode != RDG_MODE_BF16 && mode != RDG_MODE_FP16)) { rdg_set_error("rdg_critic_forward_tc: bad arguments"); return RDG_E_BADARG; }
    if (!c->critic_ready) { rdg_set_error("critic weights not set"); return RDG_E_NOWEIGHT; }
    RDG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    int r;
    if (c->critic_packed_stale) {
        for (int k = 0; k < 2; ++k)
            for (int l = 1; l < 4; ++l) {
                ConvGeom g = rdg_critic_conv_geom(c, l, 1);
                if (!c->c_wpack[k][l - 1]) RDG_CUDA(cudaMalloc(&c->c_wpack[k][l - 1], (size_t)27 * g.Ci * g.Co * 2));
                if ((r = pack_critic_weights(k == 0 ? RDG_HALF_BF16 : RDG_HALF_FP16, c->c_params + c->c_off[2 * l], c->c_wpack[k][l - 1], g.Ci, g.Co, st))) return r;
            }
        {   // first conv as a [Co][Kpad] tf32 operand image (k = tap * Ci + channel), tcg_gemm.cu
            ConvGeom g = rdg_critic_conv_geom(c, 0, 1);
            if (!c->c_w1p_score) RDG_CUDA(cudaMalloc(&c->c_w1p_score, (size_t)g.Co * tcg_smallci_kpad(27, g.Ci) * 4));
            if ((r = tcg_pack_smallci_weights(c->c_params + c->c_off[0], c->c_w1p_score, 27, g.Ci, g.Co, st))) return r;
            if (!c->c_w1q_score) RDG_CUDA(cudaMalloc(&c->c_w1q_score, (size_t)g.Co * g.Ci * 32 * 4));
            if ((r = tcg_pack_smallci_weights_chmajor(c->c_params + c->c_off[0], c->c_w1q_score, 27, g.Ci, g.Co, st))) return r;
        }
        c->launches += 7;
        c->critic_packed_stale = false;
    }
    const int hk = mode == RDG_MODE_BF16 ? RDG_HALF_BF16 : RDG_HALF_FP16, wk = hk == RDG_HALF_BF16 ? 0 : 1;
    size_t per = 0;
    for (int l = 0; l < 4; ++l) { ConvGeom g = rdg_critic_conv_geom(c, l, 1); per += (size_t)g.To * g.Ho * g.Wo * g.Co * 2 + 256; }
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((c->ws_bytes - 4096) / per, 1 << 20));
    const size_t px = (size_t)RDG_NHOURS * c->nd * c->nd;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int n = std::min(chunk, B - b0);
        uint8_t* p = reinterpret_cast<uint8_t*>(c->ws);
        auto take = [&](size_t bytes) { void* q = p; p += (bytes + 255) / 256 * 256; return q; };
        ConvGeom g0 = rdg_critic_conv_geom(c, 0, n);
        void* cur = take((size_t)n * g0.To * g0.Ho * g0.Wo * g0.Co * 2);
        if (g0.Ci <= 4 && getenv("RDG_CRITIC_D1") == nullptr) {     // tensor cores (RDG_CRITIC_D1=simt: the CUDA-core kernel)
            if ((r = tcg_critic_first_conv16(hk, sample_dev + (size_t)b0 * px, cond_dev + (size_t)b0 * c->nd * c->nd * c->ncond, c->c_w1p_score,
                                             c->c_w1q_score, c->c_params + c->c_off[1], cur, g0, st))) return r;
        } else if ((r = critic_first_conv(hk, sample_dev + (size_t)b0 * px, cond_dev + (size_t)b0 * c->nd * c->nd * c->ncond, c->c_params + c->c_off[0],
                                          c->c_params + c->c_off[1], cur, n, c->nd, c->ncond, g0, st))) return r;
        for (int l = 1; l < 4; ++l) {
            ConvGeom g = rdg_critic_conv_geom(c, l, n);
            void* y = take((size_t)n * g.To * g.Ho * g.Wo * g.Co * 2);
            if ((r = tc_critic_conv(hk, cur, c->c_wpack[wk][l - 1], c->c_params + c->c_off[2 * l + 1], y, g, c->sm_count, st))) return r;
            cur = y;
        }
        ConvGeom d = rdg_critic_dense_geom(c, n);
        if ((r = critic_dense_score(hk, cur, c->c_params + c->c_off[8], c->c_params + c->c_off[9], score_dev + b0, n, d.Ci, st))) return r;
        c->launches += 5;
    }
    return 0;
}

// ------------------------------------------------------------------ building blocks for tests
static ConvGeom geom_from(const int* g) {
    ConvGeom c{};
    c.B = g[0]; c.Ti = g[1]; c.Hi = g[2]; c.Wi = g[3]; c.Ci = g[4]; c.To = g[5]; c.Ho = g[6]; c.Wo = g[7]; c.Co = g[8];
    c.KT = g[9]; c.KH = g[10]; c.KW = g[11]; c.stride = g[12]; c.pt = g[13]; c.ph = g[14]; c.pw = g[15]; c.up = g[16];
    return c;
}
// op 0: y = act(conv(x,w)+bias); op 1: dx = conv_bwd_data(dy=a, w); op 2: dw += bwd_filter(x=a, dy=b), db += colsum
extern "C" int rdg_conv3d(int op, const int* geom17, const float* a, const float* b, const float* bias, float* out,
                          float* out2, int act, void* stream) {
    if (!geom17 || !a || !out) return RDG_E_BADARG;
    ConvGeom g = geom_from(geom17);
    cudaStream_t st = (cudaStream_t)stream;
    if (op == 0) return simt_conv_fwd(a, b, bias, out, g, act, nullptr, 1.f, st);
    if (op == 1) return simt_conv_bwd_data(a, b, out, g, st);
    if (op == 2) return simt_conv_bwd_filter(a, b, out, out2, g, st);
    // 10..12: the same three on the tensor cores (tcgen05 kind::tf32, tcg_gemm.cu); g.up selects the upsample-folded forms;
    // 13: forward with 3xTF32 operand splitting (the training mode's forward passes)
    if (op >= 10 && op <= 13) {
        const size_t wn = (size_t)g.KT * g.KH * g.KW * g.Ci * g.Co;
        const int precise = op == 13;
        if (op == 13) op = 10;
        int r = 0;
        if (!g.up && g.Ci <= 4) {          // few-channel input (critic's first conv)
            const int taps = g.KT * g.KH * g.KW;
            if (op == 10) {
                float* wp = nullptr;
                RDG_CUDA(cudaMallocAsync(&wp, (size_t)g.Co * tcg_smallci_kpad(taps, g.Ci) * 4, st));
                if (!(r = tcg_pack_smallci_weights(b, wp, taps, g.Ci, g.Co, st))) r = tcg_conv_fwd_smallci(a, wp, bias, out, g, act, nullptr, 1.f, st, nullptr, precise);
                RDG_CUDA(cudaFreeAsync(wp, st));
                return r;
            }
            if (op == 11) return simt_conv_bwd_data(a, b, out, g, st);
            if ((r = tcg_conv_bwd_filter_smallci(a, b, out, g, st))) return r;
            return out2 ? simt_colsum(b, out2, (long long)g.B * g.To * g.Ho * g.Wo, g.Co, st) : 0;
        }
        if (!g.up) {
            if (op == 10) {
                float* wT = nullptr;
                RDG_CUDA(cudaMallocAsync(&wT, wn * 4, st));
                if (!(r = tcg_transpose_blocks(b, wT, g.KT * g.KH * g.KW, g.Ci, g.Co, st))) r = tcg_conv_fwd(a, wT, bias, out, g, act, nullptr, 1.f, st, nullptr, precise);
                RDG_CUDA(cudaFreeAsync(wT, st));
                return r;
            }
            if (op == 11) return tcg_conv_bwd_data(a, b, out, g, st);
            if ((r = tcg_conv_bwd_filter(a, b, out, g, st))) return r;
            return out2 ? simt_colsum(b, out2, (long long)g.B * g.To * g.Ho * g.Wo, g.Co, st) : 0;
        }
        const size_t fn = folded_weight_elems(g.Ci, g.Co);
        float* wf = nullptr; float* wfT = nullptr;
        RDG_CUDA(cudaMallocAsync(&wf, fn * 4, st));
        RDG_CUDA(cudaMallocAsync(&wfT, fn * 4, st));
        if (op == 10) {
            if (!(r = folded_pack_f32(b, wf, g.Ci, g.Co, st)) && !(r = tcg_transpose_blocks(wf, wfT, 64, g.Ci, g.Co, st)))
                r = tcg_folded_fwd(a, wfT, bias, out, g, st, precise);
        } else if (op == 11) {   // `out` = gradient w.r.t. the LOW-RES input
            if (!(r = folded_pack_f32(b, wf, g.Ci, g.Co, st))) r = tcg_folded_bwd_data(a, wf, out, g, st);
        } else {
            RDG_CUDA(cudaMemsetAsync(wf, 0, fn * 4, st));
            if (!(r = tcg_folded_bwd_filter(a, b, wf, g, st))) r = folded_unfold_grad(wf, out, g.Ci, g.Co, st);
            if (!r && out2) r = simt_colsum(b, out2, (long long)8 * g.B * g.Ti * g.Hi * g.Wi, g.Co, st);
        }
        RDG_CUDA(cudaFreeAsync(wf, st));
        RDG_CUDA(cudaFreeAsync(wfT, st));
        return r;
    }
    return RDG_E_BADARG;
}

// One tensor-core layer in isolation: x f32 [B,T,H,W,Cin] -> rounded to 16 bit -> layer -> y f32 [B,2T,2H,2W,Cout]
extern "C" int rdg_tc_layer(rdg_ctx* c, int layer, int mode, const float* x_dev, float* y_dev, int B, void* stream) {
    if (!c || layer < 0 || layer > 2 || (mode != RDG_MODE_BF16 && mode != RDG_MODE_FP16)) return RDG_E_BADARG;
    if (!c->gen_ready) return RDG_E_NOWEIGHT;
    RDG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    static const int cin[3] = {256, 256, 128}, cout[3] = {256, 128, 64};
    const int hk = mode == RDG_MODE_BF16 ? RDG_HALF_BF16 : RDG_HALF_FP16;
    const int s = c->nd / 8, f = 1 << layer;
    const size_t n_in = (size_t)B * 3 * f * s * f * s * f * cin[layer];
    const size_t n_out = (size_t)B * 24 * f * s * f * s * f * cout[layer];
    if ((n_in + n_out) * 2 + 1024 > c->ws_bytes) { rdg_set_error("rdg_tc_layer: batch too large for workspace"); return RDG_E_BADARG; }
    uint8_t* xin = reinterpret_cast<uint8_t*>(c->ws);
    uint8_t* yout = xin + (n_in * 2 + 255) / 256 * 256;
    int r;
    if ((r = f32_to_half(hk, x_dev, xin, (long long)n_in, st))) return r;
    const int wk = hk == RDG_HALF_BF16 ? 0 : 1;
    if (layer == 2 && c->conv3_planes)
        r = tc_upconv64_planes(hk, xin, c->g_wpack_planes[wk], c->g_params + c->g_off[7], yout, nullptr, nullptr, nullptr, nullptr, B, 3 * f, c->sm_count, st);
    else
        r = tc_upconv_pixelnorm(hk, xin, c->g_wpack[wk][layer], c->g_params + c->g_off[3 + 2 * layer], yout,
                                nullptr, nullptr, B, 3 * f, s * f, s * f, cin[layer], cout[layer], c->sm_count, st);
    if (r) return r;
    return half_to_f32(hk, yout, y_dev, (long long)n_out, st);
}

extern "C" int rdg_pixelnorm(const float* x_dev, float* y_dev, long long rows, int C, int lrelu, void* stream) {
    return ew_pixelnorm(x_dev, y_dev, rows, C, lrelu, (cudaStream_t)stream);
}
extern "C" int rdg_softmax_hours(const float* logits_dev, float* out_dev, long long B, int P, void* stream) {
    return ew_softmax_hours(logits_dev, out_dev, B, P, nullptr, 1, 1, 1.f, 0, nullptr, (cudaStream_t)stream);
}
