import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rdg_oracle as O
from rdg_b200 import weights as W
from rdg_b200.engine import Context, Generator, Critic, GanTrainer
ctx = Context(16, 1, max_chunk=256)
gw = W.randomize_biases(W.init_generator_weights(0)); cw = W.randomize_biases(W.init_critic_weights(1), seed=9)
gen, crit = Generator(gw, ctx=ctx), Critic(cw, ctx=ctx)
tr = GanTrainer(gen, crit)
for seed in range(16):
    rng = np.random.default_rng(1000 + seed)
    cond = (np.clip(rng.gamma(0.8, 12.0, size=(1, 16, 16, 1)), 0, 200) / 127.4).astype(np.float32)
    z = rng.standard_normal((1, 100)).astype(np.float32)
    m = [(rng.random(s) < 0.75).astype(np.float32) for s in O.critic_mask_shapes(16, 1)]
    _, go = O.generator_step(gw, cw, z, cond, m, torch.float64)
    tr.generator_grads(z, cond, m); torch.cuda.synchronize()
    g = tr.grad_tensor(0).cpu().numpy(); off = 0; errs = []
    for i, shp in enumerate(W.generator_shapes(16, 1)):
        n = int(np.prod(shp)); a = g[off:off+n].reshape(shp); off += (n+3)//4*4
        errs.append(np.linalg.norm(a - go[i]) / max(np.linalg.norm(go[i]), 1e-30))
    print("seed", 1000 + seed, "max rel (t0..8)", f"{max(errs[:9]):.2e}", ["%.0e" % e for e in errs[:9]])
