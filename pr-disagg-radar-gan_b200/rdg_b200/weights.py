"""Keras-layout weight containers and initialisers for the RainDisaggGAN nets.

Weights are kept on the host as a flat list of float32 numpy arrays in exactly
the order/layout a Keras ``model.get_weights()`` would return them
(reference: gan_train_cwgangp_pixelnorm.py:312-357 generator, :272-309 critic):

  generator: Dense kernel (100+nd*nd*ncond, 256*(nd//8)^2*3), bias,
             Conv3D kernels (3,3,3,256,256) (3,3,3,256,128) (3,3,3,128,64) (3,3,3,64,1) + biases
  critic:    Conv3D kernels (3,3,3,1+ncond,64) (3,3,3,64,128) (3,3,3,128,256) (3,3,3,256,256) + biases,
             Dense kernel (flat,1), bias

Pure numpy; no CUDA needed.  The device-side packed forms are produced by the
C-ABI library (csrc/pack.cpp), not here.
"""
from __future__ import annotations

import numpy as np

LATENT_DIM = 100          # gan_train_cwgangp_pixelnorm.py:68
NHOURS = 24
NORM_SCALE = 127.4        # gan_train_cwgangp_pixelnorm.py:63 / raindisagg_gan_pretrained.py:13


def tf_same_pad(i: int, k: int, s: int):
    """TensorFlow padding='same': returns (out, before, after). SURVEY A3."""
    o = -(-i // s)
    p = max((o - 1) * s + k - i, 0)
    return o, p // 2, p - p // 2


def tf_valid_out(i: int, k: int, s: int) -> int:
    return (i - k) // s + 1


def critic_geometry(nd: int):
    """Per-layer (in_dims, out_dims, pads_before) for the 4 stride-2 critic convs.

    Layer 0 is padding='valid', layers 1-3 are TF 'same' (asymmetric for even
    inputs) -- gan_train_cwgangp_pixelnorm.py:286-301.
    """
    dims = (NHOURS, nd, nd)
    geo = []
    out = tuple(tf_valid_out(d, 3, 2) for d in dims)
    geo.append((dims, out, (0, 0, 0)))
    dims = out
    for _ in range(3):
        o_p = [tf_same_pad(d, 3, 2) for d in dims]
        out = tuple(o for o, _, _ in o_p)
        geo.append((dims, out, tuple(b for _, b, _ in o_p)))
        dims = out
    return geo


def generator_shapes(nd: int = 16, ncond: int = 1):
    s = nd // 8
    n_in = LATENT_DIM + nd * nd * ncond
    n_nodes = 256 * s * s * 3
    return [
        (n_in, n_nodes), (n_nodes,),
        (3, 3, 3, 256, 256), (256,),
        (3, 3, 3, 256, 128), (128,),
        (3, 3, 3, 128, 64), (64,),
        (3, 3, 3, 64, 1), (1,),
    ]


def critic_shapes(nd: int = 16, ncond: int = 1):
    geo = critic_geometry(nd)
    t, h, w = geo[-1][1]
    flat = t * h * w * 256
    return [
        (3, 3, 3, 1 + ncond, 64), (64,),
        (3, 3, 3, 64, 128), (128,),
        (3, 3, 3, 128, 256), (256,),
        (3, 3, 3, 256, 256), (256,),
        (flat, 1), (1,),
    ]


def init_generator_weights(seed: int = 0, nd: int = 16, ncond: int = 1):
    """RandomNormal(stddev=0.02) kernels, zero biases (gan_train...py:315,326-345)."""
    rng = np.random.default_rng(seed)
    out = []
    for shp in generator_shapes(nd, ncond):
        if len(shp) == 1:
            out.append(np.zeros(shp, np.float32))
        else:
            out.append((rng.standard_normal(shp) * 0.02).astype(np.float32))
    return out


def init_critic_weights(seed: int = 1, nd: int = 16, ncond: int = 1):
    """Keras default glorot_uniform kernels, zero biases (gan_train...py:286-304)."""
    rng = np.random.default_rng(seed)
    out = []
    for shp in critic_shapes(nd, ncond):
        if len(shp) == 1:
            out.append(np.zeros(shp, np.float32))
        else:
            rf = int(np.prod(shp[:-2])) if len(shp) > 2 else 1
            fan_in, fan_out = shp[-2] * rf, shp[-1] * rf
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            out.append(rng.uniform(-lim, lim, shp).astype(np.float32))
    return out


def randomize_biases(weights, seed: int = 7, scale: float = 0.05):
    """Test helper: non-zero biases so bias handling is actually exercised."""
    rng = np.random.default_rng(seed)
    return [w if w.ndim > 1 else (rng.standard_normal(w.shape) * scale).astype(np.float32)
            for w in weights]


def check_shapes(weights, shapes, what: str):
    if len(weights) != len(shapes):
        raise ValueError(f"{what}: expected {len(shapes)} tensors, got {len(weights)}")
    for i, (w, s) in enumerate(zip(weights, shapes)):
        if tuple(w.shape) != tuple(s):
            raise ValueError(f"{what}: tensor {i} has shape {tuple(w.shape)}, expected {tuple(s)}")
