// Training-step entry points (gan_train_cwgangp_pixelnorm.py:365-408, 468-482).
#include "rdg_common.cuh"
#include "gen_tc.h"
#include "ctx.h"
#include "../../include/rdg_b200.h"
#include <cmath>
#include <algorithm>

static int ensure_train_state(rdg_ctx* c) {
    if (!c->g_grads) {
        RDG_CUDA(cudaMalloc(&c->g_grads, c->g_total * 4)); RDG_CUDA(cudaMemset(c->g_grads, 0, c->g_total * 4));
        RDG_CUDA(cudaMalloc(&c->g_m, c->g_total * 4));     RDG_CUDA(cudaMemset(c->g_m, 0, c->g_total * 4));
        RDG_CUDA(cudaMalloc(&c->g_v, c->g_total * 4));     RDG_CUDA(cudaMemset(c->g_v, 0, c->g_total * 4));
        RDG_CUDA(cudaMalloc(&c->c_grads, c->c_total * 4)); RDG_CUDA(cudaMemset(c->c_grads, 0, c->c_total * 4));
        RDG_CUDA(cudaMalloc(&c->c_m, c->c_total * 4));     RDG_CUDA(cudaMemset(c->c_m, 0, c->c_total * 4));
        RDG_CUDA(cudaMalloc(&c->c_v, c->c_total * 4));     RDG_CUDA(cudaMemset(c->c_v, 0, c->c_total * 4));
    }
    return 0;
}

extern "C" int rdg_grad_buffer(rdg_ctx* c, int which, float** grads_dev, size_t* n) {
    if (!c || which < 0 || which > 1) return RDG_E_BADARG;
    RDG_CUDA(cudaSetDevice(c->device));
    int r = ensure_train_state(c);
    if (r) return r;
    if (grads_dev) *grads_dev = which == 0 ? c->g_grads : c->c_grads;
    if (n) *n = which == 0 ? c->g_total : c->c_total;
    return 0;
}
extern "C" int rdg_param_buffer(rdg_ctx* c, int which, float** params_dev, size_t* n) {
    if (!c || which < 0 || which > 1) return RDG_E_BADARG;
    if (params_dev) *params_dev = which == 0 ? c->g_params : c->c_params;
    if (n) *n = which == 0 ? c->g_total : c->c_total;
    return 0;
}

// Adam moment buffers (same flat layout as the parameters) for optimizer-state checkpoints
extern "C" int rdg_adam_buffers(rdg_ctx* c, int which, float** m_dev, float** v_dev, size_t* n) {
    if (!c || which < 0 || which > 1) return RDG_E_BADARG;
    RDG_CUDA(cudaSetDevice(c->device));
    int r = ensure_train_state(c);
    if (r) return r;
    if (m_dev) *m_dev = which == 0 ? c->g_m : c->c_m;
    if (v_dev) *v_dev = which == 0 ? c->g_v : c->c_v;
    if (n) *n = which == 0 ? c->g_total : c->c_total;
    return 0;
}

// Keras OptimizerV2 Adam, TF 2.1: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps)
extern "C" int rdg_adam_apply(rdg_ctx* c, int which, float lr, float beta1, float beta2, float eps, long long step_t,
                              float grad_scale, void* stream) {
    if (!c || which < 0 || which > 1 || step_t < 1) return RDG_E_BADARG;
    RDG_CUDA(cudaSetDevice(c->device));
    int r = ensure_train_state(c);
    if (r) return r;
    const double lr_t = (double)lr * std::sqrt(1.0 - std::pow((double)beta2, (double)step_t)) /
                        (1.0 - std::pow((double)beta1, (double)step_t));
    float* p = which == 0 ? c->g_params : c->c_params;
    float* g = which == 0 ? c->g_grads : c->c_grads;
    float* m = which == 0 ? c->g_m : c->c_m;
    float* v = which == 0 ? c->g_v : c->c_v;
    size_t n = which == 0 ? c->g_total : c->c_total;
    r = ew_adam(p, g, m, v, (long long)n, (float)lr_t, beta1, beta2, eps, grad_scale, (cudaStream_t)stream);
    if (r) return r;
    if (which == 0) { c->gen_stale_kinds = 3; c->fold32_stale = true; c->g_tcw_stale = true; }
    else { c->critic_packed_stale = true; c->c_wT_stale = true; }
    return 0;
}
extern "C" int rdg_adam_reset(rdg_ctx* c, int which) {
    if (!c || which < 0 || which > 1) return RDG_E_BADARG;
    RDG_CUDA(cudaSetDevice(c->device));
    int r = ensure_train_state(c);
    if (r) return r;
    size_t n = which == 0 ? c->g_total : c->c_total;
    RDG_CUDA(cudaMemset(which == 0 ? c->g_m : c->c_m, 0, n * 4));
    RDG_CUDA(cudaMemset(which == 0 ? c->g_v : c->c_v, 0, n * 4));
    return 0;
}

// ------------------------------------------------------------------ workspace
namespace {

struct Bump {
    uint8_t* p; uint8_t* end;
    float* f(size_t n) {
        uint8_t* r = p;
        p += (n * 4 + 255) / 256 * 256;
        return p <= end ? reinterpret_cast<float*>(r) : nullptr;
    }
};

// kind: 0 = critic step (slot 0), 1 = generator step, 2 = critic step slot 1
int ensure_train_ws(rdg_ctx* c, size_t bytes, int kind = 0) {
    void*& p = kind == 1 ? c->train_ws_gen : kind == 2 ? c->train_ws1 : c->train_ws;
    size_t& have = kind == 1 ? c->train_ws_gen_bytes : kind == 2 ? c->train_ws1_bytes : c->train_ws_bytes;
    if (have >= bytes) return 0;
    if (p) { RDG_CUDA(cudaDeviceSynchronize()); RDG_CUDA(cudaFree(p)); p = nullptr; have = 0; }
    RDG_CUDA(cudaMalloc(&p, bytes));
    have = bytes;
    return 0;
}

size_t critic_act_elems(const rdg_ctx* c, int l) {   // l = 0: input (with cond channels), 1..4: conv outputs
    if (l == 0) return (size_t)RDG_NHOURS * c->nd * c->nd * (1 + c->ncond);
    ConvGeom g = rdg_critic_conv_geom(c, l - 1, 1);
    return (size_t)g.To * g.Ho * g.Wo * g.Co;
}

// activations of one critic invocation kept for the backward passes
struct CriticActs {
    float* h[5];     // h[0] = input [B,24,nd,nd,1+ncond]; h[l] = dropout(lrelu(a[l]))
    float* a[5];     // a[l] = conv_l(h[l-1]) + b_l (pre-activation), l = 1..4
    float* score;    // [B]
};

#define TRY(x) do { int r_ = (x); if (r_) return r_; } while (0)

// Side stream for work that is independent of the backward-data chain (filter gradients, which only accumulate into the
// gradient buffer with atomics).  fork(): the side stream may start once everything queued on `main` so far has run;
// join(): `main` continues only after everything queued on the side stream.
struct SideStream {
    rdg_ctx* c; cudaStream_t main; int k;       // k: which of the context's side streams (see ctx.h)
    int init() {
        if (!c->s_aux[k]) {
            // between the critical chain (captured from a higher-priority stream, engine.py) and the prefetch streams (default,
            // lowest): the step's own join waits for these filter gradients
            int lo = 0, hi = 0;
            RDG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            int pr = -2;                                  // engine.py: chain -3, prefetch -1, generator-step forward 0
            pr = pr < hi ? hi : (pr > lo ? lo : pr);
            RDG_CUDA(cudaStreamCreateWithPriority(&c->s_aux[k], cudaStreamNonBlocking, pr));
            RDG_CUDA(cudaEventCreateWithFlags(&c->ev_fork[k], cudaEventDisableTiming));
            RDG_CUDA(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
        }
        return 0;
    }
    cudaStream_t aux() const { return c->s_aux[k]; }
    int fork() { RDG_CUDA(cudaEventRecord(c->ev_fork[k], main)); RDG_CUDA(cudaStreamWaitEvent(c->s_aux[k], c->ev_fork[k], 0)); return 0; }
    int join() { RDG_CUDA(cudaEventRecord(c->ev_join[k], c->s_aux[k])); RDG_CUDA(cudaStreamWaitEvent(main, c->ev_join[k], 0)); return 0; }
};

int critic_alloc(const rdg_ctx* c, Bump& ws, int B, CriticActs& A) {
    for (int l = 0; l < 5; ++l) {
        A.h[l] = ws.f((size_t)B * critic_act_elems(c, l));
        A.a[l] = l ? ws.f((size_t)B * critic_act_elems(c, l)) : nullptr;
        if (!A.h[l] || (l && !A.a[l])) return RDG_E_NOMEM;
    }
    A.score = ws.f(B);
    return A.score ? 0 : RDG_E_NOMEM;
}

// critic([sample, cond]) keeping pre-activations (gan_train_cwgangp_pixelnorm.py:275-304)
int critic_fwd_train(rdg_ctx* c, const float* sample, const float* cond, const float* const* masks, int B, CriticActs& A,
                     cudaStream_t st) {
    TRY(ew_critic_input(sample, cond, A.h[0], B, c->nd, c->ncond, st));
    for (int l = 0; l < 4; ++l) {
        ConvGeom g = rdg_critic_conv_geom(c, l, B);
        TRY(simt_conv_fwd(A.h[l], c->c_params + c->c_off[2 * l], c->c_params + c->c_off[2 * l + 1], A.h[l + 1], g, ACT_LRELU,
                          masks ? masks[l] : nullptr, 1.f / 0.75f, st, A.a[l + 1]));
    }
    ConvGeom d = rdg_critic_dense_geom(c, B);
    return simt_conv_fwd(A.h[4], c->c_params + c->c_off[8], c->c_params + c->c_off[9], A.score, d, ACT_NONE, nullptr, 1.f, st);
}

// Backward of sum_b dscore[b] * D(x_b) through the critic.  Accumulates weight gradients into `grads`
// (if non-null) and, if dx0 != null, returns the gradient w.r.t. the critic input h[0].
// tmp0/tmp1: scratch, each >= B * max_l act elems.
int critic_bwd(rdg_ctx* c, const CriticActs& A, const float* const* masks, const float* dscore, int B, float* grads,
               float* dx0, float* tmp0, float* tmp1, cudaStream_t st) {
    ConvGeom d = rdg_critic_dense_geom(c, B);
    SideStream ss{c, st, 0};
    TRY(ss.init());
    // filter gradients on the side stream, each next to the backward-data conv of the same layer
    if (grads) { TRY(ss.fork()); TRY(simt_conv_bwd_filter(A.h[4], dscore, grads + c->c_off[8], grads + c->c_off[9], d, ss.aux())); }
    float* dh = tmp0;   // gradient w.r.t. h[l]
    float* da = tmp1;   // gradient w.r.t. a[l]
    TRY(simt_conv_bwd_data(dscore, c->c_params + c->c_off[8], dh, d, st));
    for (int l = 4; l >= 1; --l) {
        ConvGeom g = rdg_critic_conv_geom(c, l - 1, B);
        const long long n = (long long)B * critic_act_elems(c, l);
        if (grads) TRY(ss.join());                    // the previous layer's filter gradient still reads da
        TRY(ew_lrelu_bwd(A.a[l], dh, da, n, masks ? masks[l - 1] : nullptr, 1.f / 0.75f, st));
        if (grads) {
            TRY(ss.fork());
            TRY(simt_conv_bwd_filter(A.h[l - 1], da, grads + c->c_off[2 * (l - 1)], grads + c->c_off[2 * (l - 1) + 1], g, ss.aux()));
        }
        if (l > 1) TRY(simt_conv_bwd_data(da, c->c_params + c->c_off[2 * (l - 1)], dh, g, st));
        else if (dx0) TRY(simt_conv_bwd_data(da, c->c_params + c->c_off[0], dx0, g, st));
    }
    if (grads) TRY(ss.join());
    return 0;
}

size_t gen_act(const rdg_ctx* c, int l) {   // 0 dense out, 1..3 conv outs (per sample elements)
    const int s = c->nd / 8, f = 1 << l;
    static const int ch[4] = {256, 256, 128, 64};
    return (size_t)3 * f * s * f * s * f * ch[l];
}

struct GenActs {
    float* x0;        // [B, 100 + nd*nd*ncond]
    float* d0_pre;    // dense pre-activation
    float* y[4];      // y[0] = lrelu(dense); y[l] = lrelu(pixelnorm(cpre[l]))
    float* cpre[4];   // conv outputs + bias (pre-norm), l = 1..3
    float* logits;    // [B,24,nd,nd]
    float* img;       // [B,24,nd,nd] fractions
};

int gen_alloc(const rdg_ctx* c, Bump& ws, int B, GenActs& G) {
    ConvGeom dg = rdg_gen_dense_geom(c, 1);
    G.x0 = ws.f((size_t)B * dg.Ci);
    G.d0_pre = ws.f((size_t)B * dg.Co);
    bool ok = G.x0 && G.d0_pre;
    for (int l = 0; l < 4; ++l) {
        G.y[l] = ws.f((size_t)B * gen_act(c, l));
        G.cpre[l] = l ? ws.f((size_t)B * gen_act(c, l)) : nullptr;
        ok = ok && G.y[l] && (!l || G.cpre[l]);
    }
    const size_t px = (size_t)RDG_NHOURS * c->nd * c->nd;
    G.logits = ws.f((size_t)B * px);
    G.img = ws.f((size_t)B * px);
    return ok && G.logits && G.img ? 0 : RDG_E_NOMEM;
}

// FP32 generator forward keeping everything the backward needs (gan_train_cwgangp_pixelnorm.py:319-350)
int gen_fwd_train(rdg_ctx* c, const float* latent, const float* cond, int B, GenActs& G, float* scratch, cudaStream_t st) {
    ConvGeom dg = rdg_gen_dense_geom(c, B);
    TRY(rdg_refold32(c, st));
    TRY(ew_assemble_gen_input(latent, cond, 1, 0, G.x0, B, c->nd * c->nd * c->ncond, st));
    TRY(simt_conv_fwd(G.x0, c->g_params + c->g_off[0], c->g_params + c->g_off[1], G.y[0], dg, ACT_LRELU, nullptr, 1.f, st, G.d0_pre));
    for (int l = 0; l < 3; ++l) {
        ConvGeom g = rdg_gen_conv_geom(c, l, B);
        TRY(folded_conv_fwd(G.y[l], c->g_wfold32[l], c->g_params + c->g_off[3 + 2 * l], G.cpre[l + 1], scratch, g, st));
        TRY(ew_pixelnorm(G.cpre[l + 1], G.y[l + 1], (long long)B * g.To * g.Ho * g.Wo, g.Co, 1, st));
    }
    ConvGeom g4 = rdg_gen_conv_geom(c, 3, B);
    TRY(simt_conv_fwd(G.y[3], c->g_params + c->g_off[8], c->g_params + c->g_off[9], G.logits, g4, ACT_NONE, nullptr, 1.f, st));
    return ew_softmax_hours(G.logits, G.img, B, c->nd * c->nd, nullptr, 1, 1, 1.f, 0, nullptr, st);
}

}  // namespace

namespace {
int critic_step_tc_masks(rdg_ctx* c, const float* x_real, const float* cond, const float* latent, const float* alpha,
                         const float* const* mf, const float* const* mr, const float* const* mh, int B, int gen_mode, float* losses4,
                         cudaStream_t st);
int generator_step_tc(rdg_ctx* c, const float* latent, const float* cond, const float* const* masks, int B, float* loss_dev, cudaStream_t st,
                      int phases = 3);
}

// critic_model.train_on_batch evaluation without the optimizer update
// (gan_train_cwgangp_pixelnorm.py:365-392, 472): losses4 = [total, l_valid, l_fake, l_gp]; gradients of
// `total` w.r.t. the critic weights are left in the critic gradient buffer.
extern "C" int rdg_critic_step_grads(rdg_ctx* c, const float* x_real_dev, const float* cond_dev, const float* latent_dev,
                                     const float* alpha_dev, const float* const* masks_fake, const float* const* masks_real,
                                     const float* const* masks_hat, int B, int gen_mode, float* losses4_dev, void* stream) {
    if (!c || B < 1 || !x_real_dev || !cond_dev || !latent_dev || !alpha_dev || !losses4_dev) { rdg_set_error("rdg_critic_step_grads: bad arguments"); return RDG_E_BADARG; }
    if (!c->gen_ready || !c->critic_ready) { rdg_set_error("weights not set"); return RDG_E_NOWEIGHT; }
    RDG_CUDA(cudaSetDevice(c->device));
    TRY(ensure_train_state(c));
    cudaStream_t st = (cudaStream_t)stream;
    if (c->train_mode == 1) return critic_step_tc_masks(c, x_real_dev, cond_dev, latent_dev, alpha_dev, masks_fake, masks_real, masks_hat, B, gen_mode, losses4_dev, st);
    const size_t px = (size_t)RDG_NHOURS * c->nd * c->nd;
    size_t max_act = 0, sum_act = 0;
    for (int l = 0; l < 5; ++l) { max_act = std::max(max_act, critic_act_elems(c, l)); sum_act += critic_act_elems(c, l); }
    const size_t need = ((size_t)B * (3 * (2 * sum_act + 64) + 3 * sum_act + 8 * max_act + 3 * px) + 4096) * 4 + 128 * 256;
    TRY(ensure_train_ws(c, need));
    Bump ws{reinterpret_cast<uint8_t*>(c->train_ws), reinterpret_cast<uint8_t*>(c->train_ws) + c->train_ws_bytes};
    // The three critic invocations (fake :372, real :373, interpolated :379) run as ONE batch of 3B samples [fake | real | hat]:
    // every conv sees 3x the rows (the deep critic layers have only 12 / 2 rows per sample), a third of the launches.
    CriticActs A3;
    TRY(critic_alloc(c, ws, 3 * B, A3));
    CriticActs Af = A3, Ar = A3, Ah = A3;                 // views of the three thirds
    for (int l = 0; l < 5; ++l) {
        const size_t e = (size_t)B * critic_act_elems(c, l);
        Ar.h[l] = A3.h[l] + e; Ah.h[l] = A3.h[l] + 2 * e;
        if (l) { Ar.a[l] = A3.a[l] + e; Ah.a[l] = A3.a[l] + 2 * e; }
    }
    Ar.score = A3.score + B; Ah.score = A3.score + 2 * B;
    const bool use_masks = masks_fake && masks_real && masks_hat;
    if (!use_masks && (masks_fake || masks_real || masks_hat)) { rdg_set_error("rdg_critic_step_grads: give all three mask sets or none"); return RDG_E_BADARG; }
    float* masks3_buf[4] = {nullptr, nullptr, nullptr, nullptr};
    const float* masks3[4] = {nullptr, nullptr, nullptr, nullptr};
    if (use_masks)
        for (int l = 0; l < 4; ++l) {
            const size_t e = (size_t)B * critic_act_elems(c, l + 1);
            masks3_buf[l] = ws.f(3 * e);
            if (!masks3_buf[l]) { rdg_set_error("training workspace too small"); return RDG_E_NOMEM; }
            RDG_CUDA(cudaMemcpyAsync(masks3_buf[l], masks_fake[l], e * 4, cudaMemcpyDeviceToDevice, st));
            RDG_CUDA(cudaMemcpyAsync(masks3_buf[l] + e, masks_real[l], e * 4, cudaMemcpyDeviceToDevice, st));
            RDG_CUDA(cudaMemcpyAsync(masks3_buf[l] + 2 * e, masks_hat[l], e * 4, cudaMemcpyDeviceToDevice, st));
            masks3[l] = masks3_buf[l];
        }
    float* fake_img = ws.f((size_t)B * px);
    float* xhat = ws.f((size_t)B * px);
    float* t0 = ws.f((size_t)2 * B * max_act); float* t1 = ws.f((size_t)2 * B * max_act);
    float* gbuf = ws.f((size_t)B * max_act);   // first-order gradient g_l of D(xhat)
    float* ubuf = ws.f((size_t)B * max_act);   // second-order cotangent u_l
    float* vbuf = ws.f((size_t)B * max_act);
    float* delta[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int l = 1; l <= 4; ++l) delta[l] = ws.f((size_t)B * critic_act_elems(c, l));
    float* dscore = ws.f(2 * B); float* dscore_gp = ws.f(B); float* norm = ws.f(B); float* lsc = ws.f(8);
    if (!lsc || !delta[4]) { rdg_set_error("training workspace too small"); return RDG_E_NOMEM; }

    // frozen generator forward (:370, generator.trainable = False :363)
    TRY(rdg_generator_forward(c, latent_dev, cond_dev, 1, fake_img, B, gen_mode, RDG_OUT_FRACTION, 1.f, nullptr, stream));
    RDG_CUDA(cudaMemsetAsync(c->c_grads, 0, c->c_total * 4, st));

    // three critic invocations in graph order: fake (:372), real (:373), interpolated (:379), as one 3B batch
    TRY(ew_interp(x_real_dev, fake_img, alpha_dev, xhat, B, (long long)px, st));
    TRY(ew_critic_input(fake_img, cond_dev, Af.h[0], B, c->nd, c->ncond, st));
    TRY(ew_critic_input(x_real_dev, cond_dev, Ar.h[0], B, c->nd, c->ncond, st));
    TRY(ew_critic_input(xhat, cond_dev, Ah.h[0], B, c->nd, c->ncond, st));
    for (int l = 0; l < 4; ++l) {
        ConvGeom g = rdg_critic_conv_geom(c, l, 3 * B);
        TRY(simt_conv_fwd(A3.h[l], c->c_params + c->c_off[2 * l], c->c_params + c->c_off[2 * l + 1], A3.h[l + 1], g, ACT_LRELU,
                          masks3[l], 1.f / 0.75f, st, A3.a[l + 1]));
    }
    TRY(simt_conv_fwd(A3.h[4], c->c_params + c->c_off[8], c->c_params + c->c_off[9], A3.score, rdg_critic_dense_geom(c, 3 * B), ACT_NONE,
                      nullptr, 1.f, st));

    cudaStream_t main_st = st;
    {   // the gradient-penalty chain (further down) may start as soon as the forward pass is done
        SideStream gp0{c, main_st, 1};
        TRY(gp0.init());
        TRY(gp0.fork());
    }
    // Wasserstein terms: l_valid = mean(-D(real)), l_fake = mean(+D(fake))  (:215-216, targets :452-454); their backward
    // passes run as one 2B batch [fake | real] with dscore = [+1/B | -1/B]
    TRY(ew_mean_scaled(Ar.score, B, -1.f, lsc + 0, st));
    TRY(ew_mean_scaled(Af.score, B, 1.f, lsc + 1, st));
    TRY(ew_fill(dscore, B, 1.f / (float)B, st));
    TRY(ew_fill(dscore + B, B, -1.f / (float)B, st));
    TRY(critic_bwd(c, A3, use_masks ? masks3 : nullptr, dscore, 2 * B, c->c_grads, nullptr, t0, t1, st));

    // gradient penalty (:230-244): first-order backward of sum_b D(xhat_b) to xhat, keeping delta_l.  Independent of the
    // Wasserstein backward above (separate buffers; both only ADD into the gradient buffer), so it runs on its own stream.
    SideStream gp{c, main_st, 1};
    {
        cudaStream_t st = gp.aux();               // shadows the caller's stream inside this block
        ConvGeom d = rdg_critic_dense_geom(c, B);
        TRY(ew_fill(dscore_gp, B, 1.f, st));
        TRY(simt_conv_bwd_data(dscore_gp, c->c_params + c->c_off[8], gbuf, d, st));          // g_4 = W5 broadcast
        for (int l = 4; l >= 1; --l) {
            ConvGeom g = rdg_critic_conv_geom(c, l - 1, B);
            TRY(ew_lrelu_bwd(Ah.a[l], gbuf, delta[l], (long long)B * critic_act_elems(c, l), masks_hat ? masks_hat[l - 1] : nullptr, 1.f / 0.75f, st));
            TRY(simt_conv_bwd_data(delta[l], c->c_params + c->c_off[2 * (l - 1)], gbuf, g, st));   // g_{l-1}
        }
        const int C0 = 1 + c->ncond;
        TRY(ew_gp_norm(gbuf, C0, B, (long long)px, norm, st));                            // ||grad_xhat D||_2 per sample
        TRY(ew_gp_loss(norm, B, lsc + 2, st));                                           // 'mse' against zeros
        // cotangent of 10 * mean((n-1)^2) w.r.t. g_0 (sample channel only)
        TRY(ew_gp_cotangent(gbuf, norm, 10.f * 2.f / (float)B, ubuf, C0, B, (long long)px, st));
        // second-order pass: g_{l-1} = convT_l(delta_l) is bilinear in (W_l, delta_l); LeakyReLU'' = 0
        SideStream ss{c, st, 2};
        TRY(ss.init());
        for (int l = 1; l <= 4; ++l) {
            ConvGeom g = rdg_critic_conv_geom(c, l - 1, B);
            TRY(ss.fork());                                                                                        // side stream:
            TRY(simt_conv_bwd_filter(ubuf, delta[l], c->c_grads + c->c_off[2 * (l - 1)], nullptr, g, ss.aux()));   // d/dW_l
            TRY(simt_conv_fwd(ubuf, c->c_params + c->c_off[2 * (l - 1)], nullptr, vbuf, g, ACT_NONE, nullptr, 1.f, st));   // d/d delta_l
            TRY(ss.join());                                                                                        // ubuf is rewritten next
            TRY(ew_lrelu_bwd(Ah.a[l], vbuf, ubuf, (long long)B * critic_act_elems(c, l), masks_hat ? masks_hat[l - 1] : nullptr, 1.f / 0.75f, st));
        }
        TRY(simt_colsum(ubuf, c->c_grads + c->c_off[8], B, (int)critic_act_elems(c, 4), st));                   // d/dW5
    }
    TRY(gp.join());
    TRY(ew_combine_losses(lsc + 0, lsc + 1, lsc + 2, 10.f, losses4_dev, st));
    return 0;
}

// score[b] = D(sample_b, cond_b) (optional) and grad[b] = d D(sample_b, cond_b) / d sample_b, inference mode (no dropout), FP32:
// the quantity GradientPenalty takes the norm of (gan_train_cwgangp_pixelnorm.py:238-241); critic_model.predict's third output.
extern "C" int rdg_critic_input_grad(rdg_ctx* c, const float* sample_dev, const float* cond_dev, float* score_dev, float* grad_dev, int B,
                                     void* stream) {
    if (!c || B < 1 || !sample_dev || !cond_dev || !grad_dev) { rdg_set_error("rdg_critic_input_grad: bad arguments"); return RDG_E_BADARG; }
    if (!c->critic_ready) { rdg_set_error("weights not set"); return RDG_E_NOWEIGHT; }
    RDG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t px = (size_t)RDG_NHOURS * c->nd * c->nd;
    size_t cmax = 0, csum = 0;
    for (int l = 0; l < 5; ++l) { cmax = std::max(cmax, critic_act_elems(c, l)); csum += critic_act_elems(c, l); }
    TRY(ensure_train_ws(c, ((size_t)B * (2 * csum + 3 * cmax + 64) + 4096) * 4 + 32 * 256));
    Bump ws{reinterpret_cast<uint8_t*>(c->train_ws), reinterpret_cast<uint8_t*>(c->train_ws) + c->train_ws_bytes};
    CriticActs A;
    TRY(critic_alloc(c, ws, B, A));
    float* t0 = ws.f((size_t)B * cmax); float* t1 = ws.f((size_t)B * cmax); float* dx0 = ws.f((size_t)B * cmax);
    float* dscore = ws.f(B);
    if (!dscore) { rdg_set_error("training workspace too small"); return RDG_E_NOMEM; }
    TRY(critic_fwd_train(c, sample_dev, cond_dev, nullptr, B, A, st));
    if (score_dev) RDG_CUDA(cudaMemcpyAsync(score_dev, A.score, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
    TRY(ew_fill(dscore, B, 1.f, st));
    TRY(critic_bwd(c, A, nullptr, dscore, B, nullptr, dx0, t0, t1, st));
    return ew_extract_channel0(dx0, grad_dev, (long long)B * px, 1 + c->ncond, st);
}

// generator_model.train_on_batch evaluation without the optimizer update (:395-408, :482):
// loss = mean(-D(G(z, c), c)); gradients w.r.t. the generator weights are left in the generator gradient buffer.
extern "C" int rdg_generator_step_grads(rdg_ctx* c, const float* latent_dev, const float* cond_dev, const float* const* masks,
                                        int B, float* loss_dev, void* stream) {
    if (!c || B < 1 || !latent_dev || !cond_dev || !loss_dev) { rdg_set_error("rdg_generator_step_grads: bad arguments"); return RDG_E_BADARG; }
    if (!c->gen_ready || !c->critic_ready) { rdg_set_error("weights not set"); return RDG_E_NOWEIGHT; }
    RDG_CUDA(cudaSetDevice(c->device));
    TRY(ensure_train_state(c));
    cudaStream_t st = (cudaStream_t)stream;
    if (c->train_mode == 1) return generator_step_tc(c, latent_dev, cond_dev, masks, B, loss_dev, st);
    const size_t px = (size_t)RDG_NHOURS * c->nd * c->nd;
    size_t cmax = 0, csum = 0, gsum = 0, gmax = 0;
    for (int l = 0; l < 5; ++l) { cmax = std::max(cmax, critic_act_elems(c, l)); csum += critic_act_elems(c, l); }
    for (int l = 0; l < 4; ++l) { gsum += gen_act(c, l); gmax = std::max(gmax, gen_act(c, l)); }
    ConvGeom dg1 = rdg_gen_dense_geom(c, 1);
    const size_t up_max = gmax / 64 * 128 ;   // largest upsampled-input gradient: [24,nd,nd,128]
    const size_t need = ((size_t)B * (2 * csum + 3 * cmax + 2 * gsum + dg1.Ci + dg1.Co + 4 * px + 2 * gmax + 2 * up_max + 64) + folded_weight_elems(256, 256) + 4096) * 4 + 64 * 256;
    TRY(ensure_train_ws(c, need));
    Bump ws{reinterpret_cast<uint8_t*>(c->train_ws), reinterpret_cast<uint8_t*>(c->train_ws) + c->train_ws_bytes};
    GenActs G; CriticActs A;
    TRY(gen_alloc(c, ws, B, G)); TRY(critic_alloc(c, ws, B, A));
    float* t0 = ws.f((size_t)B * cmax); float* t1 = ws.f((size_t)B * cmax);
    float* dx0 = ws.f((size_t)B * cmax);
    float* dimg = ws.f((size_t)B * px); float* dlog = ws.f((size_t)B * px);
    float* dy = ws.f((size_t)B * gmax); float* dc = ws.f((size_t)B * gmax);
    float* dup = ws.f((size_t)B * up_max);        // phase-major staging (forward) / phase-major dy (backward): >= B * gmax
    float* dwf = ws.f(folded_weight_elems(256, 256));   // folded filter gradients of one layer
    float* dscore = ws.f(B);
    if (!dscore || !dup || !dwf) { rdg_set_error("training workspace too small"); return RDG_E_NOMEM; }

    TRY(gen_fwd_train(c, latent_dev, cond_dev, B, G, dup, st));
    TRY(critic_fwd_train(c, G.img, cond_dev, masks, B, A, st));            // critic frozen (:395), dropout active
    TRY(ew_mean_scaled(A.score, B, -1.f, loss_dev, st));                   // wasserstein_loss with target -1 (:408, :452)
    TRY(ew_fill(dscore, B, -1.f / (float)B, st));
    TRY(critic_bwd(c, A, masks, dscore, B, nullptr, dx0, t0, t1, st));
    TRY(ew_extract_channel0(dx0, dimg, (long long)B * px, 1 + c->ncond, st));
    RDG_CUDA(cudaMemsetAsync(c->g_grads, 0, c->g_total * 4, st));
    TRY(ew_softmax_hours_bwd(G.img, dimg, dlog, B, c->nd * c->nd, st));
    SideStream ss{c, st, 0};
    TRY(ss.init());
    {   // output conv: filter gradient on the side stream next to the backward-data conv
        ConvGeom g4 = rdg_gen_conv_geom(c, 3, B);
        TRY(ss.fork());
        TRY(simt_conv_bwd_filter(G.y[3], dlog, c->g_grads + c->g_off[8], c->g_grads + c->g_off[9], g4, ss.aux()));
        TRY(simt_conv_bwd_data(dlog, c->g_params + c->g_off[8], dy, g4, st));
    }
    for (int l = 2; l >= 0; --l) {   // upsample + conv + pixelnorm + lrelu blocks
        ConvGeom g = rdg_gen_conv_geom(c, l, B);
        const long long rows = (long long)B * g.To * g.Ho * g.Wo;
        TRY(ss.join());                               // the previous block's filter gradient still reads dc / dup
        TRY(ew_pixelnorm_lrelu_bwd(G.cpre[l + 1], dy, dc, rows, g.Co, st));
        // upsample-folded backward: straight to the low-res input (the UpSampling3D backward is absorbed), folded filter
        // gradients (side stream) un-folded into the 3^3 kernel's gradient
        TRY(folded_deinterleave(dc, dup, g, st));
        TRY(ss.fork());
        TRY(folded_conv_bwd_filter(G.y[l], dc, dup, dwf, c->g_grads + c->g_off[2 + 2 * l], c->g_grads + c->g_off[3 + 2 * l], g, ss.aux()));
        TRY(folded_conv_bwd_data_phases(dup, c->g_wfold32[l], dy, g, st));
    }
    TRY(ss.join());
    {   // dense + lrelu
        ConvGeom dg = rdg_gen_dense_geom(c, B);
        TRY(ew_lrelu_bwd(G.d0_pre, dy, dc, (long long)B * dg.Co, nullptr, 1.f, st));
        TRY(simt_conv_bwd_filter(G.x0, dc, c->g_grads + c->g_off[0], c->g_grads + c->g_off[1], dg, st));
    }
    return 0;
}

// ================================================================================================================
// Tensor-core training mode (train_mode 1): the same two steps with every wide contraction on tcgen05 kind::tf32
// (tcg_gemm.cu) and a shorter dependency chain:
//  * the Wasserstein backward (fake | real) and the first-order gradient-penalty backward (interpolated) are ONE 3B backward-data
//    chain with cotangents [+1/B | -1/B | 1] (the same linear operator, per-sample masks);
//  * the second-order cotangents u_l overwrite the interpolated third of the saved activations h_l, so the Wasserstein filter
//    gradient bwd_filter(h_{l-1}, da_l) and the penalty's bwd_filter(u_{l-1}, delta_l) are ONE contraction over 3B samples;
//  * the generator's Conv3D(64->1) runs as per-tap products P = y . w4^T (N = 32) + a gather, its backward as two plain GEMMs
//    on the gathered cotangent matrix.
// The 2-channel first critic conv (K = 54) and the Dense(1) stay on the CUDA cores.
// ================================================================================================================
#include "tcg.h"

namespace {

int refresh_critic_wT(rdg_ctx* c, cudaStream_t st) {
    ConvGeom g0 = rdg_critic_conv_geom(c, 0, 1);
    if (!c->c_wT) RDG_CUDA(cudaMalloc(&c->c_wT, c->c_total * 4));
    if (!c->c_w1p) RDG_CUDA(cudaMalloc(&c->c_w1p, (size_t)g0.Co * tcg_smallci_kpad(27, g0.Ci) * 4));
    if (!c->c_wT_stale) return 0;
    TRY(tcg_pack_smallci_weights(c->c_params + c->c_off[0], c->c_w1p, 27, g0.Ci, g0.Co, st));
    {
        const float* src[3]; float* dst[3]; int nblk[3], R[3], C[3];
        for (int l = 1; l < 4; ++l) {
            ConvGeom g = rdg_critic_conv_geom(c, l, 1);
            src[l - 1] = c->c_params + c->c_off[2 * l]; dst[l - 1] = c->c_wT + c->c_off[2 * l];
            nblk[l - 1] = 27; R[l - 1] = g.Ci; C[l - 1] = g.Co;
        }
        TRY(tcg_transpose_blocks_batch(3, src, dst, nblk, R, C, st));
    }
    c->c_wT_stale = false;
    return 0;
}

int refresh_gen_tcw(rdg_ctx* c, cudaStream_t st) {
    static const int cin[3] = {256, 256, 128}, cout[3] = {256, 128, 64};
    TRY(rdg_refold32(c, st));
    ConvGeom dg = rdg_gen_dense_geom(c, 1);
    for (int l = 0; l < 3; ++l)
        if (!c->g_wfoldT[l]) RDG_CUDA(cudaMalloc(&c->g_wfoldT[l], folded_weight_elems(cin[l], cout[l]) * sizeof(float)));
    if (!c->g_denseT) RDG_CUDA(cudaMalloc(&c->g_denseT, (size_t)dg.Ci * dg.Co * 4));
    if (!c->g_w4p) RDG_CUDA(cudaMalloc(&c->g_w4p, 2 * 2048 * 4));
    if (!c->g_tcw_stale) return 0;
    {
        const float* src[4] = {c->g_wfold32[0], c->g_wfold32[1], c->g_wfold32[2], c->g_params + c->g_off[0]};
        float* dst[4] = {c->g_wfoldT[0], c->g_wfoldT[1], c->g_wfoldT[2], c->g_denseT};
        const int nblk[4] = {64, 64, 64, 1}, R[4] = {cin[0], cin[1], cin[2], dg.Ci}, C[4] = {cout[0], cout[1], cout[2], dg.Co};
        TRY(tcg_transpose_blocks_batch(4, src, dst, nblk, R, C, st));
    }
    TRY(ew_pad_w4(c->g_params + c->g_off[8], c->g_w4p, st));
    c->g_tcw_stale = false;
    return 0;
}

bool tc_layer_ok(const ConvGeom& g) { return (g.Ci & 31) == 0 && (g.Co & 31) == 0; }

// critic conv layer l (0..3) forward on the tensor cores; precise = 3xTF32 (real forward passes), 0 for the linear second-order pass
int critic_conv_fwd_tc(rdg_ctx* c, int l, const float* x, const float* bias, float* y, const ConvGeom& g, int act, const float* mask,
                       cudaStream_t st, float* pre, int precise) {
    if (l == 0 && g.Ci <= 4) return tcg_conv_fwd_smallci(x, c->c_w1p, bias, y, g, act, mask, 1.f / 0.75f, st, pre, precise);
    if (l && tc_layer_ok(g)) return tcg_conv_fwd(x, c->c_wT + c->c_off[2 * l], bias, y, g, act, mask, 1.f / 0.75f, st, pre, precise);
    return simt_conv_fwd(x, c->c_params + c->c_off[2 * l], bias, y, g, act, mask, 1.f / 0.75f, st, pre);
}

// masks3: per layer ONE buffer of 3B samples [fake | real | interpolated] (or all null)
// phases: bit 0 = the part that does not read the critic's weights (frozen generator forward, interpolation, critic inputs), bit 1 =
// the rest.  Issued separately, phase 1 of a step overlaps the gradient exchange + Adam update of the previous step (other stream).
int critic_step_tc(rdg_ctx* c, const float* x_real, const float* cond, const float* latent, const float* alpha, const float* const* masks3,
                   int B, int gen_mode, float* losses4, cudaStream_t st, int phases = 3, int slot = 0) {
    const size_t px = (size_t)RDG_NHOURS * c->nd * c->nd;
    size_t max_act = 0, sum_act = 0;
    for (int l = 0; l < 5; ++l) { max_act = std::max(max_act, critic_act_elems(c, l)); sum_act += critic_act_elems(c, l); }
    const size_t need = ((size_t)B * (3 * (3 * sum_act + 64) + 6 * max_act + 3 * px) + 8192) * 4 + 128 * 256;
    TRY(ensure_train_ws(c, need, slot ? 2 : 0));
    uint8_t* wsb = reinterpret_cast<uint8_t*>(slot ? c->train_ws1 : c->train_ws);
    Bump ws{wsb, wsb + (slot ? c->train_ws1_bytes : c->train_ws_bytes)};
    CriticActs A3;
    TRY(critic_alloc(c, ws, 3 * B, A3));
    float* hat_h[5]; float* hat_a[5];
    for (int l = 0; l < 5; ++l) {
        const size_t e = (size_t)B * critic_act_elems(c, l);
        hat_h[l] = A3.h[l] + 2 * e;
        hat_a[l] = l ? A3.a[l] + 2 * e : nullptr;
    }
    float* fake_img = ws.f((size_t)B * px);
    float* dh = ws.f((size_t)3 * B * max_act);
    float* da[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int l = 1; l <= 4; ++l) da[l] = ws.f((size_t)3 * B * critic_act_elems(c, l));
    float* g0 = ws.f((size_t)B * critic_act_elems(c, 0));
    float* vbuf = ws.f((size_t)B * max_act);
    float* dscore3 = ws.f(3 * B); float* norm = ws.f(B); float* lsc = ws.f(8);
    if (!lsc || !vbuf || !da[4]) { rdg_set_error("training workspace too small"); return RDG_E_NOMEM; }
    const float ms = 1.f / 0.75f;

    if (phases & 1) {
        // frozen generator forward (:370, generator.trainable = False :363)
        TRY(rdg_generator_forward(c, latent, cond, 1, fake_img, B, gen_mode, RDG_OUT_FRACTION, 1.f, nullptr, st));
        // RandomWeightedAverage (:221-224) + the three [sample, condition] concatenations (:275-282) in one pass
        TRY(ew_critic_inputs3(fake_img, x_real, alpha, cond, A3.h[0], B, c->nd, c->ncond, st));
    }
    if (!(phases & 2)) return 0;
    TcgArenaScope arena(&c->splitk_arena[0]);
    TRY(refresh_critic_wT(c, st));
    RDG_CUDA(cudaMemsetAsync(c->c_grads, 0, c->c_total * 4, st));
    // critic forward on [fake | real | interpolated] (:372, :373, :379)
    for (int l = 0; l < 4; ++l) {
        ConvGeom g = rdg_critic_conv_geom(c, l, 3 * B);
        TRY(critic_conv_fwd_tc(c, l, A3.h[l], c->c_params + c->c_off[2 * l + 1], A3.h[l + 1], g, ACT_LRELU, masks3 ? masks3[l] : nullptr, st,
                               A3.a[l + 1], 1));
    }
    // one backward-data chain for the three thirds: cotangents of the scores [+1/B | -1/B | 1].  One launch forms them, the two
    // Wasserstein terms l_fake = mean(+D(fake)) -> lsc[1], l_valid = mean(-D(real)) -> lsc[0] (:215-216, targets :452-454) and
    // da_4 = Dense(1)^T backward * LeakyReLU'(a_4) * mask_4
    // side streams: bias / Dense gradients on the first; the four filter gradients round-robin over all three, so that they run
    // next to each other and next to the second-order chain (each is a sub-wave grid) instead of queueing behind one another
    SideStream ss{c, st, 0}, ssk[3] = {{c, st, 0}, {c, st, 1}, {c, st, 2}};
    for (auto& s : ssk) TRY(s.init());
    {
        // the scores only feed the reported losses (the cotangents of the scores are constants): Dense(1) forward and the two
        // Wasserstein terms run on a side stream, the backward chain starts right away
        TRY(ss.fork());
        TRY(ew_dense_score(A3.h[4], c->c_params + c->c_off[8], c->c_params + c->c_off[9], A3.score, 3 * B, (int)critic_act_elems(c, 4), ss.aux()));
        TRY(ew_score_losses(A3.score, B, 2, 1.f, -1.f, lsc + 4, ss.aux()));       // lsc[4] = l_fake, lsc[5] = l_valid
        const float cot3[3] = {1.f / (float)B, -1.f / (float)B, 1.f}, sign2[2] = {1.f, -1.f};
        TRY(ew_critic_tail(nullptr, c->c_params + c->c_off[8], A3.a[4], masks3 ? masks3[3] : nullptr, ms, B, 3, (int)critic_act_elems(c, 4), cot3,
                           sign2, 0, nullptr, dscore3, da[4], st));
    }
    TRY(ss.fork());
    TRY(simt_conv_bwd_filter(A3.h[4], dscore3, c->c_grads + c->c_off[8], c->c_grads + c->c_off[9], rdg_critic_dense_geom(c, 2 * B), ss.aux()));
    // da_l = cotangent of the pre-activation a_l.  The LeakyReLU (+ dropout) backward of layer l-1 is fused into the epilogue of
    // layer l's transposed conv: da_{l-1} = convT_l(da_l) * LeakyReLU'(a_{l-1}) * mask_{l-1}
    // filter gradients: the Wasserstein part (fake | real, the first 2B samples: h_{l-1} and da_l are final here) starts right
    // away on a side stream and runs next to the rest of the backward chain; only the penalty part (u_{l-1}, interpolated third)
    // has to wait for the second-order pass below.  Both accumulate into the same gradient buffer.
    auto filter_grad = [&](int l, const float* x, const float* dy, int nb, cudaStream_t s) -> int {
        ConvGeom gf = rdg_critic_conv_geom(c, l - 1, nb);
        float* dw = c->c_grads + c->c_off[2 * (l - 1)];
        if (l > 1 && tc_layer_ok(gf)) return tcg_conv_bwd_filter(x, dy, dw, gf, s);
        if (l == 1 && gf.Ci <= 4) return tcg_conv_bwd_filter_smallci(x, dy, dw, gf, s);
        return simt_conv_bwd_filter(x, dy, dw, nullptr, gf, s);
    };
    for (int l = 4; l >= 1; --l) {
        ConvGeom g = rdg_critic_conv_geom(c, l - 1, 3 * B);
        TRY(ss.fork());       // bias gradients of the Wasserstein terms: the first 2B samples
        TRY(simt_colsum(da[l], c->c_grads + c->c_off[2 * (l - 1) + 1], (long long)2 * B * (critic_act_elems(c, l) / g.Co), g.Co, ss.aux()));
        {
            SideStream& sf = ssk[l % 3];
            TRY(sf.fork());
            TRY(filter_grad(l, A3.h[l - 1], da[l], 2 * B, sf.aux()));
        }
        if (l > 1) {
            if (tc_layer_ok(g)) {
                TRY(tcg_conv_bwd_data(da[l], c->c_params + c->c_off[2 * (l - 1)], da[l - 1], g, st, A3.a[l - 1], masks3 ? masks3[l - 2] : nullptr, ms));
            } else {
                TRY(simt_conv_bwd_data(da[l], c->c_params + c->c_off[2 * (l - 1)], dh, g, st));
                TRY(ew_lrelu_bwd(A3.a[l - 1], dh, da[l - 1], (long long)3 * B * critic_act_elems(c, l - 1), masks3 ? masks3[l - 2] : nullptr, ms, st));
            }
        } else {              // only the penalty needs the gradient w.r.t. the critic input: interpolated third
            ConvGeom g1 = rdg_critic_conv_geom(c, 0, B);
            TRY(simt_conv_bwd_data_ch0(da[1] + (size_t)2 * B * critic_act_elems(c, 1), c->c_params + c->c_off[0], g0, g1, st, vbuf));
        }
    }
    // gradient penalty (:230-244): norm of the input gradient, 'mse' against zeros, cotangent of 10 * mean((n-1)^2)
    const int C0 = 1 + c->ncond;
    TRY(ew_gp_fused(g0, 10.f * 2.f / (float)B, hat_h[0], C0, B, (long long)px, norm, st));      // u_0 replaces h_0 of the third
    // second-order pass (LeakyReLU'' = 0): u_l = S_l (.) conv_l(u_{l-1}); filter gradients of BOTH loss parts in one contraction
    for (int l = 1; l <= 4; ++l) {
        ConvGeom g3 = rdg_critic_conv_geom(c, l - 1, 3 * B);
        ConvGeom g = rdg_critic_conv_geom(c, l - 1, B);
        SideStream& sf = ssk[(l - 1) % 3];
        TRY(sf.fork());
        TRY(filter_grad(l, hat_h[l - 1], da[l] + (size_t)2 * B * critic_act_elems(c, l), B, sf.aux()));
        (void)g3;
        const float* mh = masks3 ? masks3[l - 1] + (size_t)2 * B * critic_act_elems(c, l) : nullptr;
        if ((l == 1 && g.Ci <= 4) || (l > 1 && tc_layer_ok(g))) {     // u_l = conv_l(u_{l-1}) * LeakyReLU'(a_l) * mask_l in one epilogue
            TRY(critic_conv_fwd_tc(c, l - 1, hat_h[l - 1], nullptr, hat_h[l], g, ACT_LRELU_BWD, mh, st, hat_a[l], 0));
        } else {
            TRY(simt_conv_fwd(hat_h[l - 1], c->c_params + c->c_off[2 * (l - 1)], nullptr, vbuf, g, ACT_NONE, nullptr, 1.f, st));
            TRY(ew_lrelu_bwd(hat_a[l], vbuf, hat_h[l], (long long)B * critic_act_elems(c, l), mh, ms, st));
        }
    }
    TRY(simt_colsum(hat_h[4], c->c_grads + c->c_off[8], B, (int)critic_act_elems(c, 4), st));      // d/dW5 of the penalty
    for (auto& s : ssk) TRY(s.join());
    TRY(ew_combine_losses_norm(lsc + 5, lsc + 4, norm, B, 10.f, losses4, st));      // [l_valid, l_fake, l_gp = mean((norm - 1)^2)]
    return 0;
}

int generator_step_tc(rdg_ctx* c, const float* latent, const float* cond, const float* const* masks, int B, float* loss_dev, cudaStream_t st,
                      int phases) {
    const size_t px = (size_t)RDG_NHOURS * c->nd * c->nd;
    size_t cmax = 0, csum = 0, gsum = 0, gmax = 0;
    for (int l = 0; l < 5; ++l) { cmax = std::max(cmax, critic_act_elems(c, l)); csum += critic_act_elems(c, l); }
    for (int l = 0; l < 4; ++l) { gsum += gen_act(c, l); gmax = std::max(gmax, gen_act(c, l)); }
    ConvGeom dg1 = rdg_gen_dense_geom(c, 1);
    const size_t need = ((size_t)B * (2 * csum + 3 * cmax + 3 * gsum + dg1.Ci + dg1.Co + 4 * px + 2 * gmax + 32 * px + 64) +
                         3 * folded_weight_elems(256, 256) + 8192) * 4 + 64 * 256;
    TRY(ensure_train_ws(c, need, 1));
    Bump ws{reinterpret_cast<uint8_t*>(c->train_ws_gen), reinterpret_cast<uint8_t*>(c->train_ws_gen) + c->train_ws_gen_bytes};
    GenActs G; CriticActs A;
    TRY(gen_alloc(c, ws, B, G)); TRY(critic_alloc(c, ws, B, A));
    float* t0 = ws.f((size_t)B * cmax); float* t1 = ws.f((size_t)B * cmax);
    float* dx0 = ws.f((size_t)B * cmax);
    float* dimg = ws.f((size_t)B * px); float* dlog = ws.f((size_t)B * px);
    float* dy = ws.f((size_t)B * gmax);
    float* dcl[3]; float* dwfl[3];                    // per block: cotangent of the conv output, folded filter gradient (read by the
    for (int l = 0; l < 3; ++l) {                     // block's own side stream while the chain moves on to the next block)
        ConvGeom g = rdg_gen_conv_geom(c, l, B);
        dcl[l] = ws.f((size_t)B * g.To * g.Ho * g.Wo * g.Co);
        dwfl[l] = ws.f(folded_weight_elems(g.Ci, g.Co));
    }
    float* ptap = ws.f((size_t)B * px * 32);          // per-tap products P (forward), gathered cotangents Gd (backward)
    float* dw4 = ws.f(2048);
    float* dscore = ws.f(B);
    if (!dscore || !dwfl[2] || !ptap) { rdg_set_error("training workspace too small"); return RDG_E_NOMEM; }
    const float ms = 1.f / 0.75f;
    ConvGeom dg = rdg_gen_dense_geom(c, B);
    ConvGeom gp{};       // the output conv's tap products as a 1x1x1 "conv" 64 -> 32 over the 24 x nd x nd grid
    gp.B = B; gp.Ti = gp.To = RDG_NHOURS; gp.Hi = gp.Ho = c->nd; gp.Wi = gp.Wo = c->nd; gp.Ci = 64; gp.Co = 32;
    gp.KT = gp.KH = gp.KW = 1; gp.stride = 1;

    if (phases & 1) {
    TcgArenaScope arena(&c->splitk_arena[1]);
    TRY(refresh_gen_tcw(c, st));
    // ---- generator forward keeping what the backward needs (:319-350); does not read the critic's weights
    TRY(ew_assemble_gen_input(latent, cond, 1, 0, G.x0, B, c->nd * c->nd * c->ncond, st));
    TRY(tcg_conv_fwd(G.x0, c->g_denseT, c->g_params + c->g_off[1], G.y[0], dg, ACT_LRELU, nullptr, 1.f, st, G.d0_pre, 1));
    for (int l = 0; l < 3; ++l) {
        ConvGeom g = rdg_gen_conv_geom(c, l, B);
        TRY(tcg_folded_fwd(G.y[l], c->g_wfoldT[l], c->g_params + c->g_off[3 + 2 * l], G.cpre[l + 1], g, st, 1));
        TRY(ew_pixelnorm(G.cpre[l + 1], G.y[l + 1], (long long)B * g.To * g.Ho * g.Wo, g.Co, 1, st));
    }
    TRY(tcg_conv_fwd(G.y[3], c->g_w4p, nullptr, ptap, gp, ACT_NONE, nullptr, 1.f, st, nullptr, 1));     // the image feeds the critic's LeakyReLUs: 3xTF32 too
    TRY(ew_tap_gather_logits(ptap, c->g_params + c->g_off[9], G.logits, B, c->nd, st));
    TRY(ew_softmax_hours(G.logits, G.img, B, c->nd * c->nd, nullptr, 1, 1, 1.f, 0, nullptr, st));
    TRY(ew_critic_input(G.img, cond, A.h[0], B, c->nd, c->ncond, st));
    }
    if (!(phases & 2)) return 0;
    TcgArenaScope arena(&c->splitk_arena[0]);
    TRY(refresh_critic_wT(c, st));

    // ---- critic on the generated sample (critic frozen :395, dropout active) and its backward to the image
    for (int l = 0; l < 4; ++l) {
        ConvGeom g = rdg_critic_conv_geom(c, l, B);
        TRY(critic_conv_fwd_tc(c, l, A.h[l], c->c_params + c->c_off[2 * l + 1], A.h[l + 1], g, ACT_LRELU, masks ? masks[l] : nullptr, st, A.a[l + 1], 1));
    }
    {   // the score only feeds the reported loss mean(-D(G(z))) (wasserstein_loss with target -1, :408, :452): off the chain
        SideStream s0{c, st, 0};
        TRY(s0.init()); TRY(s0.fork());
        TRY(ew_dense_score(A.h[4], c->c_params + c->c_off[8], c->c_params + c->c_off[9], A.score, B, (int)critic_act_elems(c, 4), s0.aux()));
        TRY(ew_score_losses(A.score, B, 1, -1.f, 0.f, loss_dev, s0.aux()));       // joined with the filter gradients below
    }
    {   // t1 / t0 alternate as the cotangents of the pre-activations a_4 .. a_1 (LeakyReLU backward fused into the transposed convs);
        // the first launch forms the cotangent -1/B of the scores and da_4
        float* cur = t1; float* nxt = t0;
        const float cot3[3] = {-1.f / (float)B, 0.f, 0.f}, sign2[2] = {-1.f, 0.f};
        TRY(ew_critic_tail(nullptr, c->c_params + c->c_off[8], A.a[4], masks ? masks[3] : nullptr, ms, B, 1, (int)critic_act_elems(c, 4), cot3, sign2,
                           0, nullptr, dscore, cur, st));
        for (int l = 4; l >= 1; --l) {
            ConvGeom g = rdg_critic_conv_geom(c, l - 1, B);
            if (l > 1 && tc_layer_ok(g)) {
                TRY(tcg_conv_bwd_data(cur, c->c_params + c->c_off[2 * (l - 1)], nxt, g, st, A.a[l - 1], masks ? masks[l - 2] : nullptr, ms));
                std::swap(cur, nxt);
            } else if (l > 1) {
                TRY(simt_conv_bwd_data(cur, c->c_params + c->c_off[2 * (l - 1)], nxt, g, st));
                TRY(ew_lrelu_bwd(A.a[l - 1], nxt, cur, (long long)B * critic_act_elems(c, l - 1), masks ? masks[l - 2] : nullptr, ms, st));
            } else {
                TRY(simt_conv_bwd_data_ch0(cur, c->c_params + c->c_off[0], dx0, g, st, nxt));
            }
        }
    }
    TRY(ew_extract_channel0(dx0, dimg, (long long)B * px, 1 + c->ncond, st));
    RDG_CUDA(cudaMemsetAsync(c->g_grads, 0, c->g_total * 4, st));
    TRY(ew_softmax_hours_bwd(G.img, dimg, dlog, B, c->nd * c->nd, st));

    // ---- generator backward
    // filter / bias gradients: the output conv's on side stream 0, block l's on side stream (3 - l) % 3 -- they only accumulate
    // into the gradient buffer and overlap each other and the backward-data chain; everything joins before the Dense layer
    SideStream ss{c, st, 0}, ssk[3] = {{c, st, 0}, {c, st, 1}, {c, st, 2}};
    for (auto& s : ssk) TRY(s.init());
    {   // output conv: Gd[pos][tap] = dlogits[pos - offset(tap)]; dy3 = Gd . w4, dw4 = Gd^T . y3, db4 = sum dlogits
        TRY(ew_tap_scatter_dlogits(dlog, ptap, B, c->nd, st));
        ConvGeom gb = gp; gb.Ci = 32; gb.Co = 64;
        TRY(ss.fork());
        RDG_CUDA(cudaMemsetAsync(dw4, 0, 2048 * 4, ss.aux()));
        TRY(tcg_conv_bwd_filter(ptap, G.y[3], dw4, gb, ss.aux()));
        RDG_CUDA(cudaMemcpyAsync(c->g_grads + c->g_off[8], dw4, 27 * 64 * 4, cudaMemcpyDeviceToDevice, ss.aux()));
        TRY(simt_colsum(dlog, c->g_grads + c->g_off[9], (long long)B * px, 1, ss.aux()));
        TRY(tcg_conv_fwd(ptap, c->g_w4p + 2048, nullptr, dy, gb, ACT_NONE, nullptr, 1.f, st));
    }
    for (int l = 2; l >= 0; --l) {   // upsample + conv + pixelnorm + lrelu blocks
        ConvGeom g = rdg_gen_conv_geom(c, l, B);
        const long long rows = (long long)B * g.To * g.Ho * g.Wo;
        float* dc = dcl[l]; float* dwf = dwfl[l];
        SideStream& sf = ssk[(3 - l) % 3];
        TRY(ew_pixelnorm_lrelu_bwd(G.cpre[l + 1], dy, dc, rows, g.Co, st));
        TRY(sf.fork());
        RDG_CUDA(cudaMemsetAsync(dwf, 0, folded_weight_elems(g.Ci, g.Co) * 4, sf.aux()));
        TRY(tcg_folded_bwd_filter(G.y[l], dc, dwf, g, sf.aux()));
        TRY(folded_unfold_grad(dwf, c->g_grads + c->g_off[2 + 2 * l], g.Ci, g.Co, sf.aux()));
        TRY(simt_colsum(dc, c->g_grads + c->g_off[3 + 2 * l], rows, g.Co, sf.aux()));
        TRY(tcg_folded_bwd_data(dc, c->g_wfold32[l], dy, g, st));
    }
    for (auto& s : ssk) TRY(s.join());
    {   // dense + lrelu
        float* dc = dcl[0];
        TRY(ew_lrelu_bwd(G.d0_pre, dy, dc, (long long)B * dg.Co, nullptr, 1.f, st));
        TRY(simt_conv_bwd_filter(G.x0, dc, c->g_grads + c->g_off[0], c->g_grads + c->g_off[1], dg, st));
    }
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------- entry points
namespace {

size_t pad4(size_t n) { return (n + 3) & ~(size_t)3; }

int ensure_rnd(rdg_ctx* c, size_t floats, int kind = 0) {     // kind as in ensure_train_ws
    float*& p = kind == 1 ? c->rnd_buf_gen : kind == 2 ? c->rnd_buf1 : c->rnd_buf;
    size_t& have = kind == 1 ? c->rnd_gen_cap : kind == 2 ? c->rnd1_cap : c->rnd_cap;
    if (have >= floats) return 0;
    if (p) { RDG_CUDA(cudaDeviceSynchronize()); RDG_CUDA(cudaFree(p)); p = nullptr; have = 0; }
    RDG_CUDA(cudaMalloc(&p, floats * 4));
    have = floats;
    return 0;
}
int ensure_tstate(rdg_ctx* c) {
    if (c->tstate) return 0;
    RDG_CUDA(cudaMalloc(&c->tstate, sizeof(RdgTrainState)));
    RDG_CUDA(cudaMemset(c->tstate, 0, sizeof(RdgTrainState)));
    return 0;
}
// layout of the per-step random inputs in rnd_buf: latent [B,100] | alpha [B] | masks of the 4 critic layers, `reps` samples sets each
struct RndLayout { size_t latent, alpha, mask[4], total; };
RndLayout rnd_layout(const rdg_ctx* c, int B, int reps) {
    RndLayout L{};
    size_t o = 0;
    L.latent = o; o += pad4((size_t)B * RDG_LATENT);
    L.alpha = o; o += pad4(B);
    for (int l = 0; l < 4; ++l) { L.mask[l] = o; o += pad4((size_t)reps * B * critic_act_elems(c, l + 1)); }
    L.total = o;
    return L;
}

}  // namespace

namespace {
int critic_step_tc_masks(rdg_ctx* c, const float* x_real, const float* cond, const float* latent, const float* alpha,
                         const float* const* mf, const float* const* mr, const float* const* mh, int B, int gen_mode, float* losses4,
                         cudaStream_t st) {
    const bool use = mf && mr && mh;
    if (!use && (mf || mr || mh)) { rdg_set_error("rdg_critic_step_grads: give all three mask sets or none"); return RDG_E_BADARG; }
    const float* masks3[4] = {nullptr, nullptr, nullptr, nullptr};
    if (use) {
        const RndLayout L = rnd_layout(c, B, 3);
        TRY(ensure_rnd(c, L.total));
        for (int l = 0; l < 4; ++l) {
            const size_t e = (size_t)B * critic_act_elems(c, l + 1);
            float* d = c->rnd_buf + L.mask[l];
            RDG_CUDA(cudaMemcpyAsync(d, mf[l], e * 4, cudaMemcpyDeviceToDevice, st));
            RDG_CUDA(cudaMemcpyAsync(d + e, mr[l], e * 4, cudaMemcpyDeviceToDevice, st));
            RDG_CUDA(cudaMemcpyAsync(d + 2 * e, mh[l], e * 4, cudaMemcpyDeviceToDevice, st));
            masks3[l] = d;
        }
    }
    return critic_step_tc(c, x_real, cond, latent, alpha, use ? masks3 : nullptr, B, gen_mode, losses4, st);
}
}  // namespace

// Derived weight images of the tensor-core training mode (transposed / folded / padded copies): rebuilt whenever the master
// weights change OUTSIDE a training step too (set_weights, checkpoint load), so that a captured iteration can be replayed on
// them without any host-side staleness check.
int rdg_refresh_train_weights(rdg_ctx* c, int which, cudaStream_t st) {
    if (which == 0) return c->gen_ready ? refresh_gen_tcw(c, st) : 0;
    return c->critic_ready ? refresh_critic_wT(c, st) : 0;
}

extern "C" int rdg_set_train_mode(rdg_ctx* c, int mode) {
    if (!c || (mode != 0 && mode != 1)) return RDG_E_BADARG;
    RDG_CUDA(cudaSetDevice(c->device));
    c->train_mode = mode;
    if (mode == 1) {
        TRY(rdg_refresh_train_weights(c, 0, nullptr));
        TRY(rdg_refresh_train_weights(c, 1, nullptr));
    }
    return 0;
}

// The two step evaluations with latent noise, interpolation weights and dropout masks drawn ON THE DEVICE from the context's
// Philox state (key = seed, counter = (element, stream, step counter in device memory)): no host work between the calls, so a
// whole 5 + 1 iteration can be captured in one CUDA graph and replayed (the step counter advances inside the graph).
extern "C" int rdg_critic_step_dev(rdg_ctx* c, const float* x_real_dev, const float* cond_dev, int B, int gen_mode,
                                   unsigned long long seed, int dropout, float* losses4_dev, int phases, void* stream) {
    if (!c || B < 1 || !x_real_dev || !cond_dev || !losses4_dev) { rdg_set_error("rdg_critic_step_dev: bad arguments"); return RDG_E_BADARG; }
    if (!c->gen_ready || !c->critic_ready) { rdg_set_error("weights not set"); return RDG_E_NOWEIGHT; }
    if (c->train_mode != 1) { rdg_set_error("rdg_critic_step_dev needs the tensor-core training mode (rdg_set_train_mode(ctx, 1))"); return RDG_E_BADARG; }
    RDG_CUDA(cudaSetDevice(c->device));
    TRY(ensure_train_state(c)); TRY(ensure_tstate(c));
    cudaStream_t st = (cudaStream_t)stream;
    const int slot = (phases & RDG_STEP_SLOT1) ? 1 : 0;
    phases &= ~RDG_STEP_SLOT1;
    if (phases < 1 || phases > 3) { rdg_set_error("rdg_critic_step_dev: phases must be 1, 2 or 3 (| RDG_STEP_SLOT1)"); return RDG_E_BADARG; }
    const RndLayout L = rnd_layout(c, B, 3);
    TRY(ensure_rnd(c, L.total, slot ? 2 : 0));
    float* rnd = slot ? c->rnd_buf1 : c->rnd_buf;
    const float* masks3[4] = {nullptr, nullptr, nullptr, nullptr};
    if (dropout)
        for (int l = 0; l < 4; ++l) masks3[l] = rnd + L.mask[l];
    if (phases & 1) {
        // latent ~ N(0,1) (np.random.normal :470), alpha ~ U[0,1) (tf.random.uniform :223) and the Dropout(0.25) keep masks of the
        // three critic invocations (:289-301: one draw over the four 3B mask tensors) in one launch that also advances the counter
        TRY(ew_step_random_dev(rnd, (long long)B * RDG_LATENT, (long long)L.alpha, B, (long long)L.mask[0],
                               dropout ? (long long)(L.total - L.mask[0]) : 0, 0.75f, seed, c->tstate, 0, st));
    }
    return critic_step_tc(c, x_real_dev, cond_dev, rnd + L.latent, rnd + L.alpha, dropout ? masks3 : nullptr, B, gen_mode,
                          losses4_dev, st, phases, slot);
}

extern "C" int rdg_generator_step_dev(rdg_ctx* c, const float* cond_dev, int B, unsigned long long seed, int dropout, float* loss_dev,
                                      int phases, void* stream) {
    if (!c || B < 1 || !cond_dev || !loss_dev) { rdg_set_error("rdg_generator_step_dev: bad arguments"); return RDG_E_BADARG; }
    if (!c->gen_ready || !c->critic_ready) { rdg_set_error("weights not set"); return RDG_E_NOWEIGHT; }
    if (c->train_mode != 1) { rdg_set_error("rdg_generator_step_dev needs the tensor-core training mode (rdg_set_train_mode(ctx, 1))"); return RDG_E_BADARG; }
    RDG_CUDA(cudaSetDevice(c->device));
    TRY(ensure_train_state(c)); TRY(ensure_tstate(c));
    cudaStream_t st = (cudaStream_t)stream;
    const RndLayout L = rnd_layout(c, B, 1);
    if (phases < 1 || phases > 3) { rdg_set_error("rdg_generator_step_dev: phases must be 1, 2 or 3"); return RDG_E_BADARG; }
    TRY(ensure_rnd(c, L.total, 1));                  // own buffer: phase 1 may run next to a critic step
    const float* masks[4] = {nullptr, nullptr, nullptr, nullptr};
    if (dropout)
        for (int l = 0; l < 4; ++l) masks[l] = c->rnd_buf_gen + L.mask[l];
    if (phases & 1) {
        // latent (generate_latent_points :177-193) and the dropout masks of the critic pass, one launch
        TRY(ew_step_random_dev(c->rnd_buf_gen, (long long)B * RDG_LATENT, (long long)L.alpha, 0, (long long)L.mask[0],
                               dropout ? (long long)(L.total - L.mask[0]) : 0, 0.75f, seed, c->tstate, 1, st));
    }
    return generator_step_tc(c, c->rnd_buf_gen + L.latent, cond_dev, dropout ? masks : nullptr, B, loss_dev, st, phases);
}

// Keras-Adam with the shared step counter in device memory (incremented by the call), followed by the refresh of every derived
// weight image the next step reads (transposed / folded / 16-bit packed), so the context is consistent at step boundaries and the
// call sequence of an iteration does not depend on host-side staleness flags (graph capture).
extern "C" int rdg_adam_apply_dev(rdg_ctx* c, int which, float lr, float beta1, float beta2, float eps, float grad_scale, int gen_mode,
                                  void* stream) {
    if (!c || which < 0 || which > 1) return RDG_E_BADARG;
    RDG_CUDA(cudaSetDevice(c->device));
    TRY(ensure_train_state(c)); TRY(ensure_tstate(c));
    cudaStream_t st = (cudaStream_t)stream;
    float* p = which == 0 ? c->g_params : c->c_params;
    float* g = which == 0 ? c->g_grads : c->c_grads;
    float* m = which == 0 ? c->g_m : c->c_m;
    float* v = which == 0 ? c->g_v : c->c_v;
    const size_t n = which == 0 ? c->g_total : c->c_total;
    TRY(ew_adam_dev(p, g, m, v, (long long)n, c->tstate, lr, beta1, beta2, eps, grad_scale, st));
    if (which == 0) {
        c->gen_stale_kinds = 3; c->fold32_stale = true; c->g_tcw_stale = true;
        if (gen_mode != RDG_MODE_FP32) TRY(rdg_repack_generator(c, st, rdg_kind_bit(gen_mode)));
        if (c->train_mode == 1) TRY(refresh_gen_tcw(c, st));
    } else {
        c->critic_packed_stale = true; c->c_wT_stale = true;
        if (c->train_mode == 1) TRY(refresh_critic_wT(c, st));
    }
    return 0;
}

// host <-> device copy of the training counters (checkpoints; keeping `optimizer.iterations` in step with replayed graphs);
// rng_ctr: two values, the Philox step counters of the critic steps and of the generator steps
extern "C" int rdg_train_state(rdg_ctx* c, int set, long long* adam_t, unsigned long long* rng_ctr) {
    if (!c || !adam_t || !rng_ctr) return RDG_E_BADARG;
    RDG_CUDA(cudaSetDevice(c->device));
    TRY(ensure_tstate(c));
    RdgTrainState s{};
    RDG_CUDA(cudaDeviceSynchronize());
    RDG_CUDA(cudaMemcpy(&s, c->tstate, sizeof(s), cudaMemcpyDeviceToHost));
    if (set) {
        s.adam_t = *adam_t; s.rng_ctr[0] = rng_ctr[0]; s.rng_ctr[1] = rng_ctr[1];
        RDG_CUDA(cudaMemcpy(c->tstate, &s, sizeof(s), cudaMemcpyHostToDevice));
    } else { *adam_t = s.adam_t; rng_ctr[0] = s.rng_ctr[0]; rng_ctr[1] = s.rng_ctr[1]; }
    return 0;
}
