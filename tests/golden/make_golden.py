"""Generates tests/golden/*.npz from the CPU oracle (FP64).  The reference itself (TensorFlow 2.1 +
the pretrained .h5) cannot run in this image, so these vectors pin the ORACLE, not TensorFlow:
they guard against silent drift of oracle and kernels alike.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200"))
import rdg_oracle as O
from rdg_b200 import weights as W

here = os.path.dirname(os.path.abspath(__file__))
gw = W.randomize_biases(W.init_generator_weights(0))
cw = W.randomize_biases(W.init_critic_weights(1), seed=9)
rng = np.random.default_rng(2024)
B = 3
cond = (np.clip(rng.gamma(0.8, 12.0, size=(B, 16, 16, 1)), 0, 200) / 127.4).astype(np.float32)
z = rng.standard_normal((B, 100)).astype(np.float32)
frac = O.generator_forward(gw, z, cond, torch.float64)
score = O.critic_forward(cw, frac.astype(np.float32), cond, None, torch.float64)
alpha = rng.random((B, 1, 1, 1, 1)).astype(np.float32)
x = rng.standard_normal((B, 24, 16, 16, 1)) * 2
x = np.exp(x - x.max(axis=1, keepdims=True)); x = (x / x.sum(axis=1, keepdims=True)).astype(np.float32)
losses, grads, _ = O.critic_step(gw, cw, x, cond, z, alpha, None, torch.float64)
np.savez_compressed(os.path.join(here, "generator_critic_nd16.npz"),
                    cond=cond, latent=z, fractions=frac.astype(np.float64), critic_score=score,
                    x_real=x, alpha=alpha, critic_losses=np.array(losses),
                    critic_grad_norms=np.array([np.linalg.norm(g) for g in grads]),
                    critic_grad_dense=grads[8])
print("wrote", os.path.join(here, "generator_critic_nd16.npz"))
