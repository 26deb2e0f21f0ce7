// Training-step entry points (gan_train_cwgangp_pixelnorm.py:365-408, 468-482).
#include "rdg_common.cuh"
#include "gen_tc.h"
#include "ctx.h"
#include "../../include/rdg_b200.h"
#include <cmath>

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
    if (which == 0) c->gen_packed_stale = true;
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

extern "C" int rdg_critic_step_grads(rdg_ctx* c, const float* x_real_dev, const float* cond_dev, const float* latent_dev,
                                     const float* alpha_dev, const float* const* masks_fake, const float* const* masks_real,
                                     const float* const* masks_hat, int B, int gen_mode, float* losses4_dev, void* stream) {
    rdg_set_error("rdg_critic_step_grads: not implemented yet");
    return RDG_E_BADARG;
}
extern "C" int rdg_generator_step_grads(rdg_ctx* c, const float* latent_dev, const float* cond_dev, const float* const* masks,
                                        int B, float* loss_dev, void* stream) {
    rdg_set_error("rdg_generator_step_grads: not implemented yet");
    return RDG_E_BADARG;
}
