#!/usr/bin/env python
"""Training-step timing (BASELINE.json configs #3 / #4): cWGAN-GP iterations/s of the gan_train_cwgangp_pixelnorm.py
step (5 critic steps + 1 generator step, batch 32 per GPU, :70-78,463-491) on synthetic radar-shaped data resident in
HBM.  Single GPU: `python tools/bench_train.py`; data-parallel: `python -m torch.distributed.run --nproc-per-node N
--master-addr 127.0.0.1 tools/bench_train.py` (one NCCL all-reduce of the flat FP32 gradient buffer per optimizer step).
Prints one JSON line (rank 0).  Secondary bench: the headline metric stays bench.py's scenarios/s."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200"))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--gen-mode", default="fp32", choices=["fp32", "fp16", "bf16"],
                    help="precision of the FROZEN generator forward inside the critic step")
    args = ap.parse_args()
    from rdg_b200 import weights as W
    from rdg_b200.engine import Context, Generator, Critic, GanTrainer
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    ctx = Context(16, 1, device=local, max_chunk=1024)
    gen = Generator(W.init_generator_weights(0), ctx=ctx)
    crit = Critic(W.init_critic_weights(1), ctx=ctx)
    tr = GanTrainer(gen, crit, gen_mode=args.gen_mode, seed=100 + rank)
    B = args.batch
    rng = np.random.default_rng(7 + rank)
    logits = rng.standard_normal((B, 24, 16, 16, 1)).astype(np.float32) * 2
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    x_real = ctx.dev((e / e.sum(axis=1, keepdims=True)).astype(np.float32))
    cond = ctx.dev((np.clip(rng.gamma(0.8, 12.0, size=(B, 16, 16, 1)), 0, 200) / 127.4).astype(np.float32))
    dev = x_real.device
    g = torch.Generator(device=dev); g.manual_seed(5 + rank)

    def iteration():
        out = None
        for _ in range(5):
            z = torch.randn((B, 100), device=dev, generator=g)
            out = tr.critic_train_on_batch([x_real, cond, z])
        z = torch.randn((B, 100), device=dev, generator=g)
        gl = tr.generator_train_on_batch([z, cond])
        return out, gl

    for _ in range(args.warmup):
        out, gl = iteration()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.iters):
        out, gl = iteration()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    dt = (time.perf_counter() - t0) / args.iters
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if dist: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "cWGAN-GP training iterations/s (5 critic + 1 generator step, batch 32 per GPU)",
                          "value": 1.0 / float(t.item()), "unit": "iterations/s", "n_gpus": world, "ms_per_iteration": 1e3 * float(t.item()),
                          "samples_per_s": world * B / float(t.item()), "scaling": "weak", "gen_mode": args.gen_mode,
                          "comm_ms_last_step": tr.comm_ms() if hasattr(tr, "comm_ms") else None,
                          "losses_last": {"critic": out, "generator": gl}}))
    if dist: dist.destroy_process_group()


if __name__ == "__main__":
    main()
