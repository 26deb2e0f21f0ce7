#!/usr/bin/env python
"""BASELINE.json config #5: the large-domain variant (alternative_domains/gan_train_cwgangp_pixelnorm_largedomain.py,
ndomain = 64: Dense 4196 -> 49152, Reshape (3,8,8,256), output 24 x 64 x 64) -- generation throughput at B = 32 / 1024 and one
critic + generator training step at batch 32, synthetic conditions, seeded random-init weights.  Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200"))
import numpy as np, torch
from rdg_b200 import weights as W
from rdg_b200.engine import Context, Critic, GanTrainer, Generator

ND = 64
MAC_DIRECT, MAC_FOLDED = 35_709_000_000, 10_845_000_000      # 71.418 / 21.690 GFLOP per scenario
ctx = Context(ND, 1)
gen = Generator(W.init_generator_weights(0, ND), ctx=ctx, mode="fp16")
crit = Critic(W.init_critic_weights(1, ND), ctx=ctx)
dev = torch.device("cuda")
rng = np.random.default_rng(3)
res = {"metric": "large-domain (nd=64) generation scenarios/s and training step", "chunk": ctx.max_chunk}
for B in (32, 1024):
    cond = ctx.dev((np.clip(rng.gamma(0.8, 12.0, size=(B, ND, ND, 1)), 0, 200) / 127.4).astype(np.float32))
    z = torch.randn((B, 100), device=dev)
    out = torch.empty((B, 24, ND, ND), device=dev)
    for _ in range(3):
        gen.forward_device(z, cond, mode="fp16", out_mm=True, out=out, check=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5 if B > 100 else 20
    e0.record()
    for _ in range(n):
        gen.forward_device(z, cond, mode="fp16", out_mm=True, out=out, check=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[f"gen_B{B}"] = {"ms": ms, "scenarios_per_s": B / ms * 1e3, "executed_tflops": B * 2 * MAC_FOLDED / ms / 1e9,
                         "direct_form_tflops": B * 2 * MAC_DIRECT / ms / 1e9}
    want = cond[:, :, :, 0] * 127.4
    if B > 100:
        import ctypes as C
        ctx.lib.rdg_profile_enable(ctx.handle, 1)
        gen.forward_device(z, cond, mode="fp16", out_mm=True, out=out, check=False)
        torch.cuda.synchronize()
        ctx.lib.rdg_profile_enable(ctx.handle, 0)
        ms_sum, n_l, n_u = (C.c_double * 8)(), (C.c_longlong * 8)(), (C.c_longlong * 8)()
        ctx.lib.rdg_profile_collect(ctx.handle, ms_sum, n_l, n_u)
        names = ["concat", "dense", "cvt16", "upconv256", "upconv128", "upconv64", "out_conv_softmax", "pixelnorm"]
        res[f"gen_B{B}"]["layer_ms"] = {names[i]: round(ms_sum[i], 3) for i in range(8) if n_l[i]}
    res[f"gen_B{B}"]["conservation_rel_err"] = float(((out.sum(dim=1) - want).abs() / want.clamp_min(1e-6)).max().item())
B = 32
tr = GanTrainer(gen, crit, gen_mode="fp16", seed=1)
lg = rng.standard_normal((B, 24, ND, ND, 1)).astype(np.float32) * 2
ex = np.exp(lg - lg.max(axis=1, keepdims=True))
x_real = ctx.dev((ex / ex.sum(axis=1, keepdims=True)).astype(np.float32))
cond = ctx.dev((np.clip(rng.gamma(0.8, 12.0, size=(B, ND, ND, 1)), 0, 200) / 127.4).astype(np.float32))
for name, fn in (("critic_step", lambda: tr.critic_train_on_batch([x_real, cond, torch.randn((B, 100), device=dev)])),
                 ("generator_step", lambda: tr.generator_train_on_batch([torch.randn((B, 100), device=dev), cond]))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        r = fn()
    torch.cuda.synchronize()
    res[name] = {"ms": (time.perf_counter() - t0) / 3 * 1e3, "batch": B, "last": r}
print(json.dumps(res))
