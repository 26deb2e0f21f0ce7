#!/usr/bin/env python
"""Where the example.py call (10 scenarios) spends its time: device time of the forward, host call overhead, copies."""
import os, sys, time, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200"))
import numpy as np, torch
from rdg_b200 import weights as W
from rdg_b200.engine import Context, Generator
ctx = Context(16, 1)
gen = Generator(W.init_generator_weights(0), ctx=ctx, mode="fp16")
B = 10
cond = np.repeat((10.0 * np.ones((1, 16, 16, 1)) / 127.4).astype(np.float32), B, axis=0)
np.random.seed(354)
lat = np.random.normal(size=(B, 100)).astype(np.float32)
zd, cd = ctx.dev(lat), ctx.dev(cond)
out = torch.empty((B, 24, 16, 16), device="cuda")
for _ in range(10): gen.forward_device(zd, cd, out=out, check=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): gen.forward_device(zd, cd, out=out, check=False)
e1.record(); torch.cuda.synchronize()
print(f"forward_device back-to-back: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call (device, launch-bound if the host is slower)")
ts = []
for _ in range(200):
    t0 = time.perf_counter(); gen.forward_device(zd, cd, out=out, check=False); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print(f"forward_device + sync: {statistics.median(ts) * 1e6:.1f} us (wall)")
t0 = time.perf_counter()
for _ in range(2000): ctx.lib.rdg_version()
print(f"ctypes call overhead: {(time.perf_counter() - t0) / 2000 * 1e6:.2f} us")
ts = []
for _ in range(300):
    t0 = time.perf_counter(); o = gen.predict([lat, cond]); ts.append(time.perf_counter() - t0)
print(f"predict (host arrays): {statistics.median(ts) * 1e6:.1f} us")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    gen.forward_device(zd, cd, out=out, check=False)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    gen.forward_device(zd, cd, out=out, check=False)
torch.cuda.synchronize()
e0.record()
for _ in range(200): g.replay()
e1.record(); torch.cuda.synchronize()
print(f"graph replay of the forward: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call (device)")
