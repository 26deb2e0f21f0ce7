"""SECOND, INDEPENDENT CPU restatement of the RainDisaggGAN graphs -- TEST INFRASTRUCTURE ONLY.

Purpose: pin oracle/rdg_oracle.py (torch ops: conv3d / interpolate / autograd) with a restatement that shares NOTHING with it:
plain numpy, explicit loops over the kernel taps, hand-written padding arithmetic, hand-written backward pass for the gradient
penalty's input gradient.  tests/test_oracle_semantics.py asserts both agree to <= 1e-12 in float64 on a B = 1 case; neither imports
the other.  Like rdg_oracle.py it is written from the reference's layer list, not from TensorFlow output (TensorFlow 2.1 and the
pretrained .h5 files are absent: PARITY UNPINNED against the reference's own arithmetic).

Reference lines followed (all in /root/reference):
  gan_train_cwgangp_pixelnorm.py:255-266  PixelNormalization.call
  gan_train_cwgangp_pixelnorm.py:272-309  create_discriminator
  gan_train_cwgangp_pixelnorm.py:312-357  create_generator
  gan_train_cwgangp_pixelnorm.py:238-241  GradientPenalty.call
  raindisagg_gan_pretrained.py:52-65      generate_scenarios
Keras / TF-2.1 layer semantics used: channels-last tensors, Conv3D = cross-correlation with kernel
(kt,kh,kw,Cin,Cout); padding='same' pads total max((ceil(i/s)-1)*s+k-i, 0) with the smaller half first; UpSampling3D = nearest
repetition; Dense kernel (in,out); Reshape / Flatten are row-major; Softmax(axis=1); LeakyReLU(alpha=0.2).
"""
import math

import numpy as np

ALPHA = 0.2


def _pad_amounts(size, k, stride, padding):
    if padding == "valid":
        return 0, 0, (size - k) // stride + 1
    out = math.ceil(size / stride)
    total = max((out - 1) * stride + k - size, 0)
    lo = total // 2
    return lo, total - lo, out


def conv3d(x, kernel, bias, stride, padding):
    """x (T,H,W,Cin) one sample; kernel (kt,kh,kw,Cin,Cout) -> (To,Ho,Wo,Cout).  One strided slice + matrix product per tap."""
    kt, kh, kw, cin, cout = kernel.shape
    T, H, W, _ = x.shape
    pt = _pad_amounts(T, kt, stride, padding)
    ph = _pad_amounts(H, kh, stride, padding)
    pw = _pad_amounts(W, kw, stride, padding)
    xp = np.zeros((T + pt[0] + pt[1], H + ph[0] + ph[1], W + pw[0] + pw[1], cin), x.dtype)
    xp[pt[0]:pt[0] + T, ph[0]:ph[0] + H, pw[0]:pw[0] + W] = x
    out = np.zeros((pt[2], ph[2], pw[2], cout), x.dtype)
    for a in range(kt):
        for b in range(kh):
            for c in range(kw):
                win = xp[a:a + (pt[2] - 1) * stride + 1:stride, b:b + (ph[2] - 1) * stride + 1:stride, c:c + (pw[2] - 1) * stride + 1:stride]
                out += win @ kernel[a, b, c]
    return out + bias


def conv3d_input_grad(dy, kernel, in_shape, stride, padding):
    """Gradient of sum(dy * conv3d(x)) w.r.t. x: every tap scatters dy @ kernel[tap]^T back onto the window it read."""
    kt, kh, kw, cin, cout = kernel.shape
    T, H, W = in_shape
    pt = _pad_amounts(T, kt, stride, padding)
    ph = _pad_amounts(H, kh, stride, padding)
    pw = _pad_amounts(W, kw, stride, padding)
    dxp = np.zeros((T + pt[0] + pt[1], H + ph[0] + ph[1], W + pw[0] + pw[1], cin), dy.dtype)
    for a in range(kt):
        for b in range(kh):
            for c in range(kw):
                dxp[a:a + (pt[2] - 1) * stride + 1:stride, b:b + (ph[2] - 1) * stride + 1:stride, c:c + (pw[2] - 1) * stride + 1:stride] += dy @ kernel[a, b, c].T
    return dxp[pt[0]:pt[0] + T, ph[0]:ph[0] + H, pw[0]:pw[0] + W]


def leaky(x):
    return np.where(x > 0, x, ALPHA * x)


def pixelnorm(x):
    return x / np.sqrt(np.mean(x * x, axis=-1, keepdims=True) + 1.0e-8)


def upsample(x):
    return np.repeat(np.repeat(np.repeat(x, 2, axis=0), 2, axis=1), 2, axis=2)


def generator_one(weights, latent, cond):
    """latent (100,), cond (nd,nd,ncond) normalised -> fractions (24,nd,nd,1).  gan_train_cwgangp_pixelnorm.py:319-350."""
    wd, bd, k1, b1, k2, b2, k3, b3, k4, b4 = weights
    nd = cond.shape[0]
    s = nd // 8
    x = np.concatenate([latent, cond.reshape(-1)])                       # Flatten + Concatenate (:321-323)
    x = leaky(x @ wd + bd).reshape(3, s, s, 256)                          # Dense, LeakyReLU, Reshape (:326-328)
    for k, b in ((k1, b1), (k2, b2), (k3, b3)):                           # (:330-343)
        x = leaky(pixelnorm(conv3d(upsample(x), k, b, 1, "same")))
    logits = conv3d(x, k4, b4, 1, "same")                                 # (:345)
    e = np.exp(logits - logits.max(axis=0, keepdims=True))                # Softmax over the hour axis (:347)
    return e / e.sum(axis=0, keepdims=True)


def critic_one(weights, sample, cond, masks=None, want_input_grad=False):
    """sample (24,nd,nd,1), cond (nd,nd,ncond) -> score (float) [, d score / d sample (24,nd,nd,1)].
    gan_train_cwgangp_pixelnorm.py:275-304; masks = the four Dropout keep masks (None: inference)."""
    k = weights[0:8:2]
    b = weights[1:8:2]
    wd, bd = weights[8], weights[9]
    x = np.concatenate([sample, np.repeat(cond[None], 24, axis=0)], axis=-1)    # (:277-282)
    pads = ("valid", "same", "same", "same")
    acts, shapes = [], []
    for l in range(4):
        shapes.append(x.shape[:3])
        a = conv3d(x, k[l], b[l], 2, pads[l])
        acts.append(a)
        x = leaky(a)
        if masks is not None:
            x = x * masks[l] / 0.75
    score = float(x.reshape(-1) @ wd[:, 0] + bd[0])
    if not want_input_grad:
        return score
    g = wd[:, 0].reshape(x.shape)
    for l in (3, 2, 1, 0):
        if masks is not None:
            g = g * masks[l] / 0.75
        g = g * np.where(acts[l] > 0, 1.0, ALPHA)
        g = conv3d_input_grad(g, k[l], shapes[l], 2, pads[l])
    return score, g[..., :1]


def gradient_penalty_term(weights, xhat, cond, masks=None):
    """(||d D(xhat) / d xhat||_2 - 1)^2 for one sample (GradientPenalty :238-241 followed by 'mse' against 0 :390)."""
    _, g = critic_one(weights, xhat, cond, masks, want_input_grad=True)
    return (math.sqrt(float((g * g).sum())) - 1.0) ** 2


def generate_scenarios_one(weights, cond_mm, latent, norm_scale=127.4):
    """raindisagg_gan_pretrained.py:52-65 for n_scenarios = 1 with the latent given: mm/h fields (24,nd,nd)."""
    cond_norm = cond_mm / norm_scale
    frac = generator_one(weights, latent, cond_norm)
    return frac[..., 0] * cond_norm[..., 0] * norm_scale
