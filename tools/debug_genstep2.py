import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from rdg_b200 import weights as W
from rdg_b200.engine import Context, Generator, Critic, GanTrainer
ctx = Context(16, 1, max_chunk=256)
gw = W.randomize_biases(W.init_generator_weights(0)); cw = W.randomize_biases(W.init_critic_weights(1), seed=9)
gen, crit = Generator(gw, ctx=ctx), Critic(cw, ctx=ctx)
rng = np.random.default_rng(21)
cond1 = (np.clip(rng.gamma(0.8, 12.0, size=(1, 16, 16, 1)), 0, 200) / 127.4).astype(np.float32)
z1 = rng.standard_normal((1, 100)).astype(np.float32)
tr = GanTrainer(gen, crit)
def grads(B):
    tr.generator_grads(np.repeat(z1, B, 0), np.repeat(cond1, B, 0), None)
    torch.cuda.synchronize()
    return tr.grad_tensor(0).cpu().numpy().copy()
g1 = grads(1)
for B in (2, 3, 4, 8, 16, 32):
    g = grads(B)
    off = 0; out = []
    for i, shp in enumerate(W.generator_shapes(16, 1)):
        n = int(np.prod(shp)); a = g[off:off+n]; b = g1[off:off+n]; off += (n + 3)//4*4
        out.append(f"{np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-30):.1e}")
    print("B", B, out[:9])
g1b = grads(1)
print("B=1 repeat diff", float(np.abs(g1b-g1).max()))
