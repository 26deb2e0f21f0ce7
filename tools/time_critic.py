#!/usr/bin/env python
"""Critic scoring throughput (samples/s): FP32 SIMT path vs the tensor-core scoring mode (test infrastructure)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200"))
import numpy as np, torch
from rdg_b200 import weights as W
from rdg_b200.engine import Context, Critic
for nd, B in ((16, 20000), (64, 1024)):
    ctx = Context(nd, 1)
    crit = Critic(W.init_critic_weights(1, nd), ctx=ctx)
    x = torch.rand((B, 24, nd, nd), device="cuda"); x = x / x.sum(dim=1, keepdim=True)
    cond = torch.rand((B, nd, nd, 1), device="cuda")
    ref = None
    for mode in ("fp32", "fp16", "bf16"):
        for _ in range(2): s = crit.forward_device(x, cond, mode=mode)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): s = crit.forward_device(x, cond, mode=mode)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        if ref is None: ref = s.clone()
        print(f"nd={nd} B={B} {mode}: {ms:.2f} ms, {B / ms * 1e3:.0f} samples/s, max |score - fp32| = {float((s - ref).abs().max()):.2e} (spread {float(ref.abs().max()):.2e})")
    ctx.close()
