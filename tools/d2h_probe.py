#!/usr/bin/env python
"""Pinned device->host copy bandwidth with all ranks copying at once (the ceiling of bench.py's `e2e` leg, which moves 24.6 GB of
fields per step and GPU).  `python tools/d2h_probe.py` (1 GPU) or `python -m torch.distributed.run --nproc-per-node N
--master-addr 127.0.0.1 tools/d2h_probe.py`.  Prints one JSON line on rank 0: per-rank min / max and aggregate GB/s."""
import json, os, sys
import torch

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pr-disagg-radar-gan_b200"))
from rdg_b200.dist import bind_to_gpu_numa_node
numa = bind_to_gpu_numa_node(local) if world > 1 else None
n = 1 << 28                                       # 1 GiB of float32
src = torch.empty(n, device="cuda", dtype=torch.float32).normal_()
dst = torch.empty(n, dtype=torch.float32, pin_memory=True)
res = {}
for name, a, b in (("d2h", src, dst), ("h2d", dst, src)):
    for _ in range(2):
        b.copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    if dist: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        b.copy_(a, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = torch.tensor([5 * n * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9], device="cuda", dtype=torch.float64)
    if dist:
        allv = [torch.zeros_like(gbs) for _ in range(world)]
        dist.all_gather(allv, gbs)
        vals = [float(v.item()) for v in allv]
    else:
        vals = [float(gbs.item())]
    res[name] = {"per_rank_min_GBps": round(min(vals), 1), "per_rank_max_GBps": round(max(vals), 1), "aggregate_GBps": round(sum(vals), 1)}
if rank == 0:
    print(json.dumps({"probe": "pinned copies, 1 GiB x 5, all ranks concurrently", "n_gpus": world, "numa_node_rank0": numa, **res}))
if dist: dist.destroy_process_group()
