"""CPU ORACLE for the RainDisaggGAN hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, with plain torch-CPU ops, the Keras graph that the reference
delegates to TensorFlow 2.1 (tensorflow=2.1.0 / cudnn=7.6.5, pr-disagg-env.yml:21-22,
118-121; TensorFlow itself is NOT under /root/reference and is not installed here).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it; the product path (pr-disagg-radar-gan_b200/) never does.

PARITY UNPINNED: the reference ships no tests, golden vectors or stored generator
outputs for this path and neither TensorFlow nor the
pretrained .h5 files exist in this container, so this restatement cannot be
checked against the reference's own arithmetic.  What pins it instead:
hand-computed known-answer cases (tests/test_oracle_semantics.py), the
conservation invariant visible in the reference's shipped CSVs, and FP64-vs-FP32
self-consistency.

Reference lines followed (all in /root/reference):
  gan_train_cwgangp_pixelnorm.py:215-216  wasserstein_loss
  gan_train_cwgangp_pixelnorm.py:219-227  RandomWeightedAverage
  gan_train_cwgangp_pixelnorm.py:230-244  GradientPenalty
  gan_train_cwgangp_pixelnorm.py:249-270  PixelNormalization
  gan_train_cwgangp_pixelnorm.py:272-309  create_discriminator
  gan_train_cwgangp_pixelnorm.py:312-357  create_generator
  gan_train_cwgangp_pixelnorm.py:360-408  combined models / losses / Adam
  gan_train_cwgangp_pixelnorm.py:463-491  train loop ordering
  raindisagg_gan_pretrained.py:52-65      generate_scenarios
  alternative_domains/gan_train_cwgangp_pixelnorm_largedomain.py:323-335  large domain
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

NHOURS = 24
LATENT_DIM = 100
NORM_SCALE = 127.4
GP_WEIGHT = 10.0  # gan_train_cwgangp_pixelnorm.py:66


# --------------------------------------------------------------------------- helpers
def tf_same_pad(i, k, s):
    """TF padding='same' (SURVEY A3): out, pad_before, pad_after."""
    o = -(-i // s)
    p = max((o - 1) * s + k - i, 0)
    return o, p // 2, p - p // 2


def _t(x, dtype):
    return torch.as_tensor(np.asarray(x)).to(dtype)


def _conv3d_keras(x, k, b, stride, padding):
    """x: (B,T,H,W,C) channels-last; k: Keras (kt,kh,kw,Cin,Cout); cross-correlation.

    padding: 'same' | 'valid' (TF semantics: 'same' pads asymmetrically, extra at the end).
    """
    xin = x.permute(0, 4, 1, 2, 3)
    if padding == "same":
        pads = []
        for d in (3, 2, 1):  # F.pad wants last dim first: W, H, T
            _, pb, pa = tf_same_pad(x.shape[d], 3, stride)
            pads += [pb, pa]
        xin = F.pad(xin, pads)
    w = k.permute(4, 3, 0, 1, 2)
    y = F.conv3d(xin, w, b, stride=stride)
    return y.permute(0, 2, 3, 4, 1)


def pixel_norm(x):
    """gan_train_cwgangp_pixelnorm.py:255-266: x / sqrt(mean_c(x^2) + 1e-8)."""
    return x / torch.sqrt(torch.mean(x * x, dim=-1, keepdim=True) + 1.0e-8)


def lrelu(x):
    return torch.where(x > 0, x, 0.2 * x)


def upsample2(x):
    """UpSampling3D((2,2,2)) nearest, channels-last."""
    return x.repeat_interleave(2, 1).repeat_interleave(2, 2).repeat_interleave(2, 3)


# --------------------------------------------------------------------------- generator
def generator_forward(weights, latent, cond, dtype=torch.float32, return_logits=False):
    """create_generator() graph (gan_train...py:312-357), inference mode.

    latent (B,100), cond (B,nd,nd,ncond) already divided by norm_scale.
    Returns fractions (B,24,nd,nd,1) as numpy in `dtype`.
    """
    w = [_t(a, dtype) for a in weights]
    z = _t(latent, dtype)
    c = _t(cond, dtype)
    B, nd = c.shape[0], c.shape[1]
    s = nd // 8
    x = torch.cat([z, c.reshape(B, -1)], dim=1)            # :321-323
    x = lrelu(x @ w[0] + w[1]).reshape(B, 3, s, s, 256)    # :326-328 (largedomain :323-335)
    for i in (2, 4, 6):                                    # :330-343
        x = upsample2(x)
        x = _conv3d_keras(x, w[i], w[i + 1], 1, "same")
        x = lrelu(pixel_norm(x))
    logits = _conv3d_keras(x, w[8], w[9], 1, "same")       # :345
    out = torch.softmax(logits, dim=1)                     # :347 softmax over hours
    if not torch.isfinite(out).all():                      # :349-350 check_numerics
        raise FloatingPointError("found nan in output of per_gridpoint_softmax")
    if return_logits:
        return out.numpy(), logits.numpy()
    return out.numpy()


def fold_upsample_conv(k):
    """Fold nearest-x2 upsample + 3^3 'same' conv into 8 phase-specific 2^3 convs.

    k: torch (3,3,3,Cin,Cout).  Returns (8 phases, 8 taps, Cin, Cout); phase index
    = pt*4+ph*2+pw, tap index = at*4+ah*2+aw, low-res offset d = a - 1 + p per axis.
    """
    Fm = torch.zeros(2, 2, 3, dtype=k.dtype)
    Fm[0, 0, 0] = 1; Fm[0, 1, 1] = 1; Fm[0, 1, 2] = 1      # even: d=-1 <- w0 ; d=0 <- w1+w2
    Fm[1, 0, 0] = 1; Fm[1, 0, 1] = 1; Fm[1, 1, 2] = 1      # odd : d=0 <- w0+w1 ; d=+1 <- w2
    f = torch.einsum("pax,qby,rcz,xyzio->pqrabcio", Fm, Fm, Fm, k)
    return f.reshape(8, 8, k.shape[3], k.shape[4])


def _folded_layer(x, kf, b):
    """x (B,T,H,W,Cin) low-res; kf (8,8,Cin,Cout) -> (B,2T,2H,2W,Cout) pre-activation."""
    B, T, H, W, _ = x.shape
    Co = kf.shape[-1]
    xp = F.pad(x, (0, 0, 1, 1, 1, 1, 1, 1))
    out = x.new_zeros(B, 2 * T, 2 * H, 2 * W, Co)
    for p in range(8):
        pt, ph, pw = p >> 2, (p >> 1) & 1, p & 1
        acc = x.new_zeros(B, T, H, W, Co)
        for a in range(8):
            at, ah, aw = a >> 2, (a >> 1) & 1, a & 1
            dt, dh, dw = at - 1 + pt, ah - 1 + ph, aw - 1 + pw
            xs = xp[:, 1 + dt:1 + dt + T, 1 + dh:1 + dh + H, 1 + dw:1 + dw + W]
            acc = acc + xs @ kf[p, a]
        out[:, pt::2, ph::2, pw::2] = acc + b
    return out


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def generator_forward_folded(weights, latent, cond, dtype=torch.float32, emulate_bf16=False,
                             return_logits=False):
    """Same function computed with the upsample fold; optional emulation of the BF16
    tensor-core mode's rounding points (dense inputs and kernel rounded to bf16, fp32
    accumulate, bf16 activations; folded weights rounded once to bf16; fp32 accumulate,
    PixelNorm/LeakyReLU in fp32, activations stored bf16; last conv with bf16 weights on the
    bf16 activations, fp32 accumulate; softmax in fp32).
    """
    w = [_t(a, dtype) for a in weights]
    z = _t(latent, dtype)
    c = _t(cond, dtype)
    B, nd = c.shape[0], c.shape[1]
    s = nd // 8
    q = _bf16 if emulate_bf16 else (lambda t: t)
    x = torch.cat([z, c.reshape(B, -1)], dim=1)
    x = q(lrelu(q(x) @ q(w[0]) + w[1])).reshape(B, 3, s, s, 256)
    for i in (2, 4, 6):
        kf = q(fold_upsample_conv(w[i]))
        x = q(lrelu(pixel_norm(_folded_layer(x, kf, w[i + 1]))))
    logits = _conv3d_keras(x, q(w[8]), w[9], 1, "same")
    out = torch.softmax(logits, dim=1)
    if return_logits:
        return out.numpy(), logits.numpy()
    return out.numpy()


def generate_scenarios(weights, cond, n_scenarios, dtype=torch.float32):
    """raindisagg_gan_pretrained.py:52-65, including its numpy-global-RNG draw and dtype quirks."""
    cond = cond / NORM_SCALE
    latent = np.random.normal(size=(n_scenarios, LATENT_DIM))
    cond_batch = np.repeat(cond[np.newaxis], repeats=n_scenarios, axis=0)
    generated = generator_forward(weights, latent, cond_batch, dtype)
    generated = generated.squeeze()
    return generated * cond.squeeze() * NORM_SCALE


# --------------------------------------------------------------------------- critic
def _critic_graph(w, sample, cond, masks=None):
    """create_discriminator() (gan_train...py:272-309) on torch tensors.

    masks: None (inference: Dropout is identity) or list of 4 {0,1} tensors shaped like
    each conv's output; kept values are scaled by 1/0.75.
    """
    B = sample.shape[0]
    ncond = cond.shape[-1]
    c = cond.reshape(B, 1, cond.shape[1], cond.shape[2], ncond).expand(-1, NHOURS, -1, -1, -1)  # :277-280
    x = torch.cat([sample, c], dim=-1)                                                        # :282
    pads = ["valid", "same", "same", "same"]
    for li in range(4):                                                                       # :286-301
        x = lrelu(_conv3d_keras(x, w[2 * li], w[2 * li + 1], 2, pads[li]))
        if masks is not None:
            x = x * masks[li] * (1.0 / 0.75)
    x = x.reshape(B, -1)                                                                      # :303
    return x @ w[8] + w[9]                                                                    # :304


def critic_forward(weights, sample, cond, masks=None, dtype=torch.float32):
    w = [_t(a, dtype) for a in weights]
    m = None if masks is None else [_t(a, dtype) for a in masks]
    return _critic_graph(w, _t(sample, dtype), _t(cond, dtype), m).numpy()


def critic_mask_shapes(nd, B):
    dims = (NHOURS, nd, nd)
    shapes = []
    out = tuple((d - 3) // 2 + 1 for d in dims)
    chans = (64, 128, 256, 256)
    shapes.append((B,) + out + (chans[0],))
    for li in range(1, 4):
        out = tuple(tf_same_pad(d, 3, 2)[0] for d in out)
        shapes.append((B,) + out + (chans[li],))
    return shapes


def _generator_graph(w, z, c):
    B, nd = c.shape[0], c.shape[1]
    s = nd // 8
    x = torch.cat([z, c.reshape(B, -1)], dim=1)
    x = lrelu(x @ w[0] + w[1]).reshape(B, 3, s, s, 256)
    for i in (2, 4, 6):
        x = upsample2(x)
        x = _conv3d_keras(x, w[i], w[i + 1], 1, "same")
        x = lrelu(pixel_norm(x))
    return torch.softmax(_conv3d_keras(x, w[8], w[9], 1, "same"), dim=1)


def critic_step(gen_w, crit_w, x_real, cond, latent, alpha, masks3, dtype=torch.float32):
    """One critic_model.train_on_batch evaluation (gan_train...py:365-392, 472) WITHOUT the
    optimizer update: returns ([total, l_valid, l_fake, l_gp], grads wrt critic weights,
    extras dict).  masks3 = three lists of 4 dropout masks, for the (fake, real, interpolated)
    critic invocations in graph-construction order (:372, :373, :379); None = no dropout.
    alpha: (B,1,1,1,1) in [0,1).
    """
    gw = [_t(a, dtype) for a in gen_w]
    cw = [_t(a, dtype).requires_grad_(True) for a in crit_w]
    xr, c, z, al = (_t(a, dtype) for a in (x_real, cond, latent, alpha))
    mk = [None, None, None] if masks3 is None else [[_t(m, dtype) for m in ms] for ms in masks3]
    with torch.no_grad():
        fake_img = _generator_graph(gw, z, c)                  # generator frozen (:363)
    fake = _critic_graph(cw, fake_img, c, mk[0])
    valid = _critic_graph(cw, xr, c, mk[1])
    xhat = (al * xr + (1 - al) * fake_img).requires_grad_(True)  # :221-224
    d_hat = _critic_graph(cw, xhat, c, mk[2])
    (g,) = torch.autograd.grad(d_hat.sum(), xhat, create_graph=True)   # K.gradients :240
    gp = torch.sqrt(torch.sum(g.reshape(g.shape[0], -1) ** 2, dim=1, keepdim=True)) - 1   # :241
    l_valid = torch.mean(-1.0 * valid)                         # targets :452-454, loss :215-216
    l_fake = torch.mean(1.0 * fake)
    l_gp = torch.mean((gp - 0.0) ** 2)                         # 'mse' vs dummy zeros
    total = l_valid + l_fake + GP_WEIGHT * l_gp                # loss_weights [1,1,10] :392
    grads = torch.autograd.grad(total, cw)
    losses = [float(total.detach()), float(l_valid.detach()), float(l_fake.detach()), float(l_gp.detach())]
    extras = dict(fake_img=fake_img.numpy(), fake=fake.detach().numpy(), valid=valid.detach().numpy(),
                  gp=gp.detach().numpy(), xhat_grad=g.detach().numpy())
    return losses, [gr.numpy() for gr in grads], extras


def generator_step(gen_w, crit_w, latent, cond, masks=None, dtype=torch.float32):
    """generator_model.train_on_batch evaluation (gan_train...py:395-408, 482) without the
    optimizer update: returns (g_loss, grads wrt generator weights).  Critic in training
    mode (dropout masks apply) but frozen."""
    gw = [_t(a, dtype).requires_grad_(True) for a in gen_w]
    cw = [_t(a, dtype) for a in crit_w]
    z, c = _t(latent, dtype), _t(cond, dtype)
    mk = None if masks is None else [_t(m, dtype) for m in masks]
    img = _generator_graph(gw, z, c)
    valid = _critic_graph(cw, img, c, mk)
    loss = torch.mean(-1.0 * valid)
    grads = torch.autograd.grad(loss, gw)
    return float(loss.detach()), [g.numpy() for g in grads]


def adam_update(params, grads, v_state, t, lr=1e-4, beta1=0.0, beta2=0.9, eps=1e-7, m_state=None):
    """Keras OptimizerV2 Adam, TF 2.1.  t is the ALREADY-INCREMENTED shared step
    counter.  Returns (new_params, new_v[, new_m]).  float64 scalar maths for lr_t like TF's
    python-side computation in float32 would differ by <1ulp; we keep float32 tensors."""
    lr_t = lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    new_p, new_v, new_m = [], [], []
    for i, (p, g, v) in enumerate(zip(params, grads, v_state)):
        p = np.asarray(p, np.float32); g = np.asarray(g, np.float32); v = np.asarray(v, np.float32)
        m_prev = np.zeros_like(p) if m_state is None else m_state[i]
        m = np.float32(beta1) * m_prev + np.float32(1 - beta1) * g
        v = np.float32(beta2) * v + np.float32(1 - beta2) * g * g
        p = p - np.float32(lr_t) * m / (np.sqrt(v) + np.float32(eps))
        new_p.append(p.astype(np.float32)); new_v.append(v.astype(np.float32)); new_m.append(m)
    return new_p, new_v, new_m


# ---------------------------------------------------------------------------------------------------------
# ensemble statistics: host reductions that follow gen.predict in the reference
# ---------------------------------------------------------------------------------------------------------
def area_mean(fields):
    """np.mean(generated * cond * norm_scale, (2, 3)) (generate_and_evaluate.py:533-535): fields [B,24,ny,nx] -> [B,24]."""
    return np.asarray(fields, np.float64).mean(axis=(2, 3))


def crps_ensemble(obs, forecasts):
    """properscoring.crps_ensemble(obs, forecasts, axis=0) with equal weights, restated from its published algorithm
    (properscoring 0.1, `_crps_ensemble_gufunc`: third-party dependency of generate_and_evaluate_crps.py:189, not in
    /root/reference and not installed here): sort the members, integrate (F_ens(x) - 1{x >= obs})^2 dx piecewise.
    obs [...], forecasts [n, ...] -> [...] (float64)."""
    obs = np.asarray(obs, np.float64)
    f = np.sort(np.asarray(forecasts, np.float64), axis=0)
    n = f.shape[0]
    flat_o, flat_f = obs.reshape(-1), f.reshape(n, -1)
    res = np.empty_like(flat_o)
    for p in range(flat_o.size):
        o, integral, obs_cdf, fc_cdf, prev = flat_o[p], 0.0, 0.0, 0.0, 0.0
        for k in range(n):
            x = flat_f[k, p]
            if obs_cdf == 0 and o < x:
                integral += (o - prev) * fc_cdf ** 2
                integral += (x - o) * (fc_cdf - 1) ** 2
                obs_cdf = 1.0
            else:
                integral += (x - prev) * (fc_cdf - obs_cdf) ** 2
            fc_cdf += 1.0 / n
            prev = x
        if obs_cdf == 0:
            integral += o - prev
        res[p] = integral
    return res.reshape(obs.shape)


def crps_ensemble_energy(obs, forecasts):
    """Same score in the energy form mean|x - y| - 0.5 mean|x - x'| (vectorised; what the CUDA kernel evaluates)."""
    f = np.asarray(forecasts, np.float64)
    o = np.asarray(obs, np.float64)
    a = np.abs(f - o[None]).mean(axis=0)
    fs = np.sort(f, axis=0)
    n = f.shape[0]
    k = (2 * np.arange(1, n + 1) - n - 1).reshape((n,) + (1,) * (f.ndim - 1))
    return a - (k * fs).sum(axis=0) / (n * n)


# ---------------------------------------------------------------------------------------------------------
# batch sampler
# ---------------------------------------------------------------------------------------------------------
def sample_windows(data, idcs_batch, ndomain, norm_scale=127.4, with_batch=True):
    """The numpy statements of generate_real_samples (gan_train_cwgangp_pixelnorm.py:148-163) for given index rows
    (tidx, yidx, xidx): window gather (view_as_windows :151-152), daily sum (:156), fractions (:159-160), normalised
    condition (:163).  float32 throughout, as in the reference (data.dtype == float32, :138)."""
    data = np.asarray(data)
    assert data.dtype == np.float32
    batch = np.stack([data[t, :, y:y + ndomain, x:x + ndomain] for t, y, x in np.asarray(idcs_batch)])
    batch = np.expand_dims(batch, -1)
    batch_cond = np.sum(batch, axis=1)
    if with_batch:
        with np.errstate(invalid="ignore", divide="ignore"):
            for i in range(len(batch)):
                batch[i] = batch[i] / batch_cond[i]
    batch_cond = batch_cond / norm_scale
    return (batch if with_batch else None), batch_cond
