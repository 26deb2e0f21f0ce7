"""GPU debug harness: each tensor-core layer in isolation vs the oracle's folded-layer emulation,
then the whole forward in all modes.  Usage (GPU box): python tools/debug_tc.py"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rdg_oracle as O
from rdg_b200 import _lib, weights as W
from rdg_b200.engine import Context, Generator

torch.set_num_threads(os.cpu_count())
ctx = Context(16, 1, max_chunk=512)
gw = W.randomize_biases(W.init_generator_weights(0))
gen = Generator(gw, ctx=ctx)
lib = ctx.lib
print("sm_count", ctx.sm_count, flush=True)

CIN, COUT = [256, 256, 128], [256, 128, 64]
for mode, name, tdt in ((_lib.MODE_FP16, "fp16", torch.float16), (_lib.MODE_BF16, "bf16", torch.bfloat16)):
    for layer in (2, 1, 0):
        f = 1 << layer
        for B in (2, 37):
            rng = np.random.default_rng(layer * 10 + B)
            x = rng.standard_normal((B, 3 * f, 2 * f, 2 * f, CIN[layer])).astype(np.float32)
            xq = torch.as_tensor(x).to(tdt).float()
            kf = O.fold_upsample_conv(torch.as_tensor(gw[2 + 2 * layer])).to(tdt).float()
            ref = O.lrelu(O.pixel_norm(O._folded_layer(xq, kf, torch.as_tensor(gw[3 + 2 * layer])))).numpy()
            xd = ctx.dev(x)
            yd = torch.empty(ref.shape, device=xd.device, dtype=torch.float32)
            t0 = time.time()
            rc = lib.rdg_tc_layer(ctx.handle, layer, mode, C.c_void_p(xd.data_ptr()), C.c_void_p(yd.data_ptr()), B, None)
            torch.cuda.synchronize()
            _lib.check(rc)
            y = yd.cpu().numpy()
            err = np.abs(y - ref)
            print(f"[{name}] layer {layer} B={B}: max abs err {err.max():.4e} mean {err.mean():.3e} "
                  f"(ref absmax {np.abs(ref).max():.3f}) nan={np.isnan(y).sum()} t={time.time()-t0:.3f}s", flush=True)
            if err.max() > 0.05:
                # structure of the error: per phase, per batch, per t
                for p in range(8):
                    pt, ph, pw = p >> 2, (p >> 1) & 1, p & 1
                    e = err[:, pt::2, ph::2, pw::2]
                    print(f"   phase {p}: max {e.max():.3e} mean {e.mean():.3e}")
                print("   per-b max", err.reshape(B, -1).max(1)[:8])
                print("   per-t max", err.max(axis=(0, 2, 3, 4)))
                print("   per-c max (first 16)", err.max(axis=(0, 1, 2, 3))[:16])
                print("   sample y", y[0, 0, 0, 0, :8], "ref", ref[0, 0, 0, 0, :8])

rng = np.random.default_rng(354)
B = 24
cond = (np.clip(rng.gamma(0.8, 12.0, size=(B, 16, 16, 1)), 0, 200) / 127.4).astype(np.float32)
z = rng.standard_normal((B, 100)).astype(np.float32)
ref = O.generator_forward(gw, z, cond, torch.float64)
for mode in ("fp32", "fp16", "bf16"):
    out = gen.predict([z, cond], mode=mode)
    rel = np.abs(out - ref) / np.abs(ref)
    print(f"full forward {mode}: max rel {rel.max():.3e} mean rel {rel.mean():.3e} cons {np.abs(out.sum(1)-1).max():.2e}", flush=True)

# quick timing of the device path
for mode in ("fp16", "bf16", "fp32"):
    Bt = 4736 if mode != "fp32" else 512
    ctx2 = Context(16, 1, max_chunk=4736) if mode == "fp16" else ctx2
    if mode == "fp16":
        gen2 = Generator(gw, ctx=ctx2)
    zc = torch.randn(Bt, 100, device="cuda")
    cc = torch.rand(Bt, 16, 16, 1, device="cuda")
    out = torch.empty(Bt, 24, 16, 16, device="cuda")
    for _ in range(2):
        gen2.forward_device(zc, cc, mode=mode, out=out, check=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        gen2.forward_device(zc, cc, mode=mode, out=out, check=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"timing {mode}: B={Bt} {ms:.3f} ms -> {Bt/ms*1e3:.0f} scenarios/s", flush=True)
