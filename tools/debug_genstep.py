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
for B in (8, 1, 32):
    cond = (np.clip(rng.gamma(0.8, 12.0, size=(B, 16, 16, 1)), 0, 200) / 127.4).astype(np.float32)
    z = rng.standard_normal((B, 100)).astype(np.float32)
    ref_loss, ref_grads = O.generator_step(gw, cw, z, cond, None, torch.float64)
    tr = GanTrainer(gen, crit)
    loss = float(tr.generator_grads(z, cond, None).item())
    g = tr.grad_tensor(0).cpu().numpy()
    print("B", B, "loss", loss, ref_loss)
    off = 0
    for i, (rg, shp) in enumerate(zip(ref_grads, W.generator_shapes(16, 1))):
        n = int(np.prod(shp)); mine = g[off:off + n].reshape(shp); off += (n + 3) // 4 * 4
        rel = np.linalg.norm(mine - rg) / np.linalg.norm(rg)
        print(f"  tensor {i} {shp}: rel l2 {rel:.3e}  max|ref| {np.abs(rg).max():.3e}")
        if i == 0:
            e = np.abs(mine - rg)
            print("    latent rows rel", np.linalg.norm((mine - rg)[:100]) / np.linalg.norm(rg[:100]), "cond rows rel", np.linalg.norm((mine - rg)[100:]) / np.linalg.norm(rg[100:]))
            print("    err by col block of 256:", [f"{np.linalg.norm((mine-rg)[:, k*256:(k+1)*256])/np.linalg.norm(rg[:, k*256:(k+1)*256]):.1e}" for k in range(12)])
