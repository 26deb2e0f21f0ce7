#!/usr/bin/env python
"""Training-step timing (BASELINE.json configs #3 / #4): one cWGAN-GP iteration of gan_train_cwgangp_pixelnorm.py
(5 critic steps + 1 generator step, batch 32 per GPU, :70-78, :463-491) on synthetic radar-shaped data resident in HBM.

  keras   : critic_model.train_on_batch / generator_model.train_on_batch surface (losses read back every step)
  device  : rdg_*_step_dev calls issued eagerly (noise / alpha / masks drawn on the GPU, nothing read back)
  graph   : the same call sequence captured once in a CUDA graph and replayed

Single GPU: `python tools/bench_train.py`; data-parallel: `python -m torch.distributed.run --nproc-per-node N
--master-addr 127.0.0.1 tools/bench_train.py` (gradient all-reduce per optimizer step, inside the graph).
Prints one JSON line (rank 0).  Secondary bench: the headline metric stays bench.py's scenarios/s."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200"))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--nd", type=int, default=16)
    ap.add_argument("--gen-mode", default="fp16", choices=["fp32", "fp16", "bf16"],
                    help="precision of the FROZEN generator forward inside the critic step")
    ap.add_argument("--paths", default="keras_fp32,keras_tf32,device,graph")
    ap.add_argument("--profile-steps", action="store_true", help="CUDA-event time of one critic step and one generator step (device path)")
    ap.add_argument("--segmented", action="store_true", help="graph path: per-phase graphs (the data-parallel form) on one GPU too")
    ap.add_argument("--dump-after", type=float, default=0, help="debug: dump all Python stacks after this many seconds")
    args = ap.parse_args()
    if args.dump_after > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.dump_after, exit=True)
    from rdg_b200 import weights as W
    from rdg_b200.engine import Context, Generator, Critic, GanTrainer
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    nd = args.nd
    ctx = Context(nd, 1, device=local, max_chunk=1024 if nd == 16 else 64)
    gen = Generator(W.init_generator_weights(0, nd), ctx=ctx, mode=args.gen_mode)
    crit = Critic(W.init_critic_weights(1, nd), ctx=ctx)
    B = args.batch
    rng = np.random.default_rng(7 + rank)
    logits = rng.standard_normal((5, B, 24, nd, nd, 1)).astype(np.float32) * 2
    e = np.exp(logits - logits.max(axis=2, keepdims=True))
    x_real = ctx.dev((e / e.sum(axis=2, keepdims=True)).astype(np.float32))
    cond = ctx.dev((np.clip(rng.gamma(0.8, 12.0, size=(5, B, nd, nd, 1)), 0, 200) / 127.4).astype(np.float32))
    dev = x_real.device
    g = torch.Generator(device=dev); g.manual_seed(5 + rank)

    def timed(fn, iters):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        if dist: dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / iters
        if dist: dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / iters, wall * 1e3], device=dev, dtype=torch.float64)
        if dist: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item())

    res = {}
    paths = args.paths.split(",")
    for name in paths:
        print(f"[rank {rank}] path {name}", file=sys.stderr, flush=True)
        if name.startswith("keras"):
            tr = GanTrainer(gen, crit, gen_mode=args.gen_mode, seed=100, train_mode=name.split("_")[1])

            def iteration():
                for k in range(5):
                    z = torch.randn((B, 100), device=dev, generator=g)
                    out = tr.critic_train_on_batch([x_real[k], cond[k], z])
                z = torch.randn((B, 100), device=dev, generator=g)
                return out, tr.generator_train_on_batch([z, cond[0]])
            res[name] = timed(iteration, max(3, args.iters // 4))
        elif name == "device":
            tr = GanTrainer(gen, crit, gen_mode=args.gen_mode, seed=100, train_mode="tf32")
            dl = torch.zeros((5, 4), device=dev); gl = torch.zeros(1, device=dev)

            def iteration():
                tr.iteration_device(x_real, cond, cond[0], dl, gl)
                tr.finish()
            res[name] = timed(iteration, args.iters)
            if args.profile_steps:
                res["critic_step_ms"] = timed(lambda: (tr.critic_step_device(x_real[0], cond[0], dl[0]), tr.finish()), args.iters * 5)[0]
                res["generator_step_ms"] = timed(lambda: (tr.generator_step_device(cond[0], gl), tr.finish()), args.iters * 5)[0]
            res["losses_last"] = {"critic": dl[-1].cpu().tolist(), "generator": float(gl.item())}
        elif name == "graph":
            tr = GanTrainer(gen, crit, gen_mode=args.gen_mode, seed=100, train_mode="tf32")
            tr.profile_comm = False
            ig = tr.capture_iteration(B, segmented=True if args.segmented else None)
            ig.x_real.copy_(x_real); ig.cond.copy_(cond); ig.cond_gen.copy_(cond[0])
            res[name] = timed(ig.replay, args.iters)
            res["graph_losses_last"] = {"critic": ig.d_losses[-1].cpu().tolist(), "generator": float(ig.g_loss.item())}
    if rank == 0:
        best = min(v[0] for k, v in res.items() if isinstance(v, tuple))
        print(json.dumps({"metric": "cWGAN-GP training iteration (5 critic + 1 generator step, batch %d per GPU, nd %d)" % (B, nd),
                          "ms_per_iteration": {k: {"device_ms": v[0], "wall_ms": v[1]} for k, v in res.items() if isinstance(v, tuple)},
                          "best_ms": best, "samples_per_s": world * B / (best * 1e-3), "n_gpus": world, "scaling": "weak",
                          "gen_mode": args.gen_mode, **{k: v for k, v in res.items() if not isinstance(v, tuple)}}))
    if dist: dist.destroy_process_group()


if __name__ == "__main__":
    main()
