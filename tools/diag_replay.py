#!/usr/bin/env python
"""Determinism diagnostic: eager device steps vs themselves vs a replayed graph (same seeds and counters)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200"))
import numpy as np, torch
from rdg_b200 import weights as W
from rdg_b200.engine import Adam, Context, Critic, GanTrainer, Generator

ctx = Context(16, 1, max_chunk=1024)
B = 8
rng = np.random.default_rng(31)
x = rng.standard_normal((5, B, 24, 16, 16, 1)) * 2
x = np.exp(x - x.max(axis=2, keepdims=True)); x = (x / x.sum(axis=2, keepdims=True)).astype(np.float32)
cond = (np.clip(rng.gamma(0.8, 12.0, size=(5, B, 16, 16, 1)), 0, 200) / 127.4).astype(np.float32)
xr, cd = torch.as_tensor(x, device="cuda"), torch.as_tensor(cond, device="cuda")

def fresh():
    g = Generator(W.init_generator_weights(7), ctx=ctx, mode="fp16")
    c = Critic(W.init_critic_weights(8), ctx=ctx)
    ctx.lib.rdg_adam_reset(ctx.handle, 0); ctx.lib.rdg_adam_reset(ctx.handle, 1)
    return GanTrainer(g, c, Adam(1e-4, 0.0, 0.9), gen_mode="fp16", train_mode="tf32", seed=5)

def snap(tr, which):
    tr.finish(); torch.cuda.synchronize()
    return tr.grad_tensor(which).cpu().numpy().copy(), tr.param_tensor(which).cpu().numpy().copy()

def eager(n_crit=1, gen=True):
    tr = fresh()
    dl = torch.zeros((5, 4), device="cuda"); gl = torch.zeros(1, device="cuda")
    out = []
    for k in range(n_crit):
        tr.critic_step_device(xr[k], cd[k], dl[k])
        out.append(snap(tr, 1))
    if gen:
        tr.generator_step_device(cd[0], gl)
        out.append(snap(tr, 0))
    return out

def cmp(a, b, name):
    for i, ((ga, pa), (gb, pb)) in enumerate(zip(a, b)):
        rg = np.linalg.norm(ga - gb) / (np.linalg.norm(gb) + 1e-30)
        fp = np.mean(np.abs(pa - pb) <= 1e-7)
        print(f"{name} step {i}: grad rel diff {rg:.3e}, params identical fraction {fp:.4f}, max param diff {np.abs(pa - pb).max():.2e}", flush=True)

a, b = eager(2), eager(2)
cmp(a, b, "eager vs eager")
