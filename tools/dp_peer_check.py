#!/usr/bin/env python
"""Data-parallel exchange over NVLink peer memory (csrc/dp_peer.cu) checked against NCCL, under torchrun on >= 2 GPUs of one node:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_peer_check.py

1. rdg_peer_allreduce of rank-dependent random gradients == dist.all_reduce of the same buffers (same values on every rank);
2. a captured data-parallel iteration (ONE graph, exchange inside) keeps the replicas bit-identical over several replays;
3. time per iteration: that graph vs the NCCL form (one graph per step phase, eager all-reduce in between).
Prints one JSON line on rank 0; exit code 1 if a check fails."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pr-disagg-radar-gan_b200"))
import numpy as np
import torch
import torch.distributed as dist


def main():
    import ctypes as C
    from rdg_b200 import _lib, weights as W
    from rdg_b200.engine import Context, Critic, GanTrainer, Generator
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")
    iters = int(os.environ.get("DP_CHECK_ITERS", "30"))
    out, ok = {"ranks": world}, True

    def build(peer):
        os.environ["RDG_PEER_ALLREDUCE"] = "1" if peer else "0"
        ctx = Context(16, 1, device=local, max_chunk=1024)
        gen = Generator(W.init_generator_weights(0), ctx=ctx, mode="fp16")
        crit = Critic(W.init_critic_weights(1), ctx=ctx)
        return ctx, GanTrainer(gen, crit, gen_mode="fp16", seed=100, train_mode="tf32")

    ctx, tr = build(True)
    out["peer_exchange"] = bool(tr.peer_exchange)
    if not tr.peer_exchange:
        ok = False
    else:
        # 1. values
        g = torch.Generator(device=dev); g.manual_seed(11 + rank)
        errs = {}
        for which, name in ((1, "critic"), (0, "generator")):
            buf = tr.grad_tensor(which)
            buf.copy_(torch.randn(buf.shape, device=dev, generator=g))
            ref = buf.clone()
            dist.all_reduce(ref)
            torch.cuda.synchronize()
            _lib.check(ctx.lib.rdg_peer_allreduce(ctx.handle, which, ctx._stream()))
            torch.cuda.synchronize()
            errs[name] = float(((buf - ref).abs().max() / ref.abs().max()).item())
            same = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(same, buf)
            errs[name + "_ranks_identical"] = bool(all(torch.equal(same[0], s) for s in same))
            ok &= errs[name] <= 1e-6 and errs[name + "_ranks_identical"]
        out["allreduce_vs_nccl_max_rel"] = errs

    B = 32
    rng = np.random.default_rng(7 + rank)
    lg = rng.standard_normal((5, B, 24, 16, 16, 1)).astype(np.float32) * 2
    ex = np.exp(lg - lg.max(axis=2, keepdims=True))
    x_real = (ex / ex.sum(axis=2, keepdims=True)).astype(np.float32)
    cond = (np.clip(rng.gamma(0.8, 12.0, size=(5, B, 16, 16, 1)), 0, 200) / 127.4).astype(np.float32)

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([a.elapsed_time(b) / n], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def run(trainer, context, label):
        nonlocal ok
        ig = trainer.capture_iteration(B)
        ig.x_real.copy_(context.dev(x_real)); ig.cond.copy_(context.dev(cond)); ig.cond_gen.copy_(context.dev(cond[0]))
        ms = timed(ig.replay, iters)
        trainer.finish(); torch.cuda.synchronize()
        # replicas: every rank must hold the same weights after the same number of updates
        ident = True
        for which in (0, 1):
            p = trainer.param_tensor(which)
            allp = [torch.empty_like(p) for _ in range(world)]
            dist.all_gather(allp, p)
            ident &= all(torch.equal(allp[0], q) for q in allp)
        finite = bool(torch.isfinite(ig.d_losses).all().item() and torch.isfinite(ig.g_loss).all().item())
        out[label] = {"ms_per_iteration": ms, "one_graph": ig.graph is not None, "replicas_bit_identical": bool(ident), "losses_finite": finite}
        ok &= ident and finite

    if tr.peer_exchange:
        run(tr, ctx, "peer")
        w, t = C.c_int(0), C.c_int(0)
        _lib.check(ctx.lib.rdg_peer_status(ctx.handle, C.byref(w), C.byref(t)))
        out["peer"]["barrier_timeouts"] = int(t.value)
        ok &= t.value == 0
    ctx2, tr2 = build(False)
    run(tr2, ctx2, "nccl")
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["ok"] = bool(flag.item())
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
