#!/usr/bin/env python
"""Per-tensor relative L2 error of the tensor-core training mode's gradients against the FP64 oracle (diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import rdg_oracle as O
from rdg_b200 import weights as W
from rdg_b200.engine import Context, Critic, Generator, GanTrainer

def rel(a, b): return float(np.linalg.norm(a.astype(np.float64) - b) / (np.linalg.norm(b) + 1e-30))
def split(flat, shapes):
    out, off = [], 0
    for s in shapes:
        n = int(np.prod(s)); out.append(flat[off:off + n].reshape(s)); off += (n + 3) // 4 * 4
    return out

ctx = Context(16, 1, max_chunk=1024)
gw = W.randomize_biases(W.init_generator_weights(0)); cw = W.randomize_biases(W.init_critic_weights(1), seed=9)
gen, crit = Generator(gw, ctx=ctx), Critic(cw, ctx=ctx)
B = 32
rng = np.random.default_rng(11)
x = rng.standard_normal((B, 24, 16, 16, 1)) * 2
x = np.exp(x - x.max(axis=1, keepdims=True)); x = (x / x.sum(axis=1, keepdims=True)).astype(np.float32)
cond = (np.clip(rng.gamma(0.8, 12.0, size=(B, 16, 16, 1)), 0, 200) / 127.4).astype(np.float32)
z = rng.standard_normal((B, 100)).astype(np.float32)
alpha = rng.random((B, 1, 1, 1, 1)).astype(np.float32)
masks3 = [[(rng.random(s) < 0.75).astype(np.float32) for s in O.critic_mask_shapes(16, B)] for _ in range(3)]
for use in (False, True):
    m3 = masks3 if use else None
    rl, rg, _ = O.critic_step(gw, cw, x, cond, z, alpha, m3, torch.float64)
    for mode in ("fp32", "tf32"):
        tr = GanTrainer(gen, crit, gen_mode="fp32", train_mode=mode)
        l = tr.critic_grads(x, cond, z, alpha.reshape(-1), m3).cpu().numpy()
        g = split(tr.grad_tensor(1).cpu().numpy(), W.critic_shapes(16, 1))
        print(f"critic masks={use} {mode}: losses err {np.abs(l - rl).max():.2e}; grads " + " ".join(f"{rel(a, b):.1e}" for a, b in zip(g, rg)), flush=True)
for Bg in (1, 8, 32):
    rng = np.random.default_rng(40 + Bg)
    cond = (np.clip(rng.gamma(0.8, 12.0, size=(Bg, 16, 16, 1)), 0, 200) / 127.4).astype(np.float32)
    z = rng.standard_normal((Bg, 100)).astype(np.float32)
    masks = [(rng.random(s) < 0.75).astype(np.float32) for s in O.critic_mask_shapes(16, Bg)]
    rl, rg = O.generator_step(gw, cw, z, cond, masks, torch.float64)
    for mode in ("fp32", "tf32"):
        tr = GanTrainer(gen, crit, train_mode=mode)
        l = float(tr.generator_grads(z, cond, masks).item())
        g = split(tr.grad_tensor(0).cpu().numpy(), W.generator_shapes(16, 1))
        print(f"generator B={Bg} {mode}: loss err {abs(l - rl):.2e}; grads " + " ".join(f"{rel(a, b):.1e}" for a, b in zip(g[:9], rg[:9])), flush=True)
