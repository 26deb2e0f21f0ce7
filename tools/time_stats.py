#!/usr/bin/env python
"""Time the ensemble-statistics kernels on device-resident fields and the fused host call (test infrastructure)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200"))
import numpy as np, torch
from rdg_b200 import weights as W
from rdg_b200.engine import Context, Generator
ctx = Context(16, 1); gen = Generator(W.init_generator_weights(0), ctx=ctx)
for spc in (100, 1000):
    n_cond = 4700 // spc
    f = torch.rand((n_cond * spc, 24, 16, 16), device="cuda"); obs = torch.rand((n_cond, 24, 16, 16), device="cuda")
    for what in ("area", "crps"):
        for _ in range(2): gen.ensemble_stats_device(f, spc, obs if what == "crps" else None)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): gen.ensemble_stats_device(f, spc, obs if what == "crps" else None)
        e1.record(); torch.cuda.synchronize()
        print(f"spc {spc} {what}: {e0.elapsed_time(e1) / 5:.3f} ms per {n_cond * spc} scenarios")
n_cond, spc = 1000, 100
rng = np.random.default_rng(0)
z = torch.randn((n_cond * spc, 100)).pin_memory(); cond = torch.rand((n_cond, 16, 16, 1)).pin_memory(); obs = torch.rand((n_cond, 24, 16, 16)).pin_memory()
for o in (None, obs):
    gen.generate_ensemble_stats_host(z, cond, spc, o)
    t0 = time.perf_counter(); gen.generate_ensemble_stats_host(z, cond, spc, o); dt = time.perf_counter() - t0
    print(f"generate_stats_host obs={'yes' if o is not None else 'no'}: {dt * 1e3:.1f} ms per {n_cond * spc} scenarios = {n_cond * spc / dt:.0f}/s")
out = torch.empty((n_cond * spc, 24, 16, 16)).pin_memory()
gen.generate_ensemble_host(z, cond, spc, out=out)
t0 = time.perf_counter(); gen.generate_ensemble_host(z, cond, spc, out=out); dt = time.perf_counter() - t0
print(f"generate_host: {dt * 1e3:.1f} ms = {n_cond * spc / dt:.0f}/s")
