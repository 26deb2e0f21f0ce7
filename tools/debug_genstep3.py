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
rng = np.random.default_rng(21)
B = 4
cond = (np.clip(rng.gamma(0.8, 12.0, size=(B, 16, 16, 1)), 0, 200) / 127.4).astype(np.float32)
z = rng.standard_normal((B, 100)).astype(np.float32)
tr = GanTrainer(gen, crit)
def mine(zz, cc):
    tr.generator_grads(zz, cc, None); torch.cuda.synchronize()
    return tr.grad_tensor(0).cpu().numpy().copy()
def orac(zz, cc):
    _, gr = O.generator_step(gw, cw, zz, cc, None, torch.float64)
    return gr
gm = mine(z, cond); gm_avg = sum(mine(z[i:i+1], cond[i:i+1]) for i in range(B)) / B
go = orac(z, cond); go_each = [orac(z[i:i+1], cond[i:i+1]) for i in range(B)]
off = 0
for i, shp in enumerate(W.generator_shapes(16, 1)):
    n = int(np.prod(shp)); a = gm[off:off+n].reshape(shp); b = gm_avg[off:off+n].reshape(shp); off += (n+3)//4*4
    oavg = sum(g[i] for g in go_each) / B
    r = lambda x, y: np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-30)
    print(f"tensor {i}: mine(B) vs mine avg {r(a,b):.1e} | oracle(B) vs oracle avg {r(go[i],oavg):.1e} | mine(B) vs oracle(B) {r(a,go[i]):.1e} | mine avg vs oracle avg {r(b,oavg):.1e}")
