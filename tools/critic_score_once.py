#!/usr/bin/env python
"""One tensor-core critic scoring pass at nd = 16, B = 20000 (ncu target for tc_critic_conv_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200"))
import torch
from rdg_b200 import weights as W
from rdg_b200.engine import Context, Critic
ctx = Context(16, 1)
crit = Critic(W.init_critic_weights(1, 16), ctx=ctx)
B = 20000
x = torch.rand((B, 24, 16, 16), device="cuda"); x = x / x.sum(dim=1, keepdim=True)
cond = torch.rand((B, 16, 16, 1), device="cuda")
for _ in range(2):
    s = crit.forward_device(x, cond, mode="fp16")
torch.cuda.synchronize()
print(float(s.sum()))
