"""Multi-GPU plumbing (torch.distributed; NCCL on GPUs, gloo in the CPU tests).

Scenario generation shards by contiguous ranges with no data-path collective;
data-parallel training exchanges one flat FP32 gradient buffer per optimizer step.
"""
from __future__ import annotations


def shard_range(total: int, rank: int, world: int, align: int = 1):
    """Contiguous [lo, hi) slice of `total` units for `rank`; slices differ by at most `align` units
    and, when align > 1 (e.g. scenarios per condition), start on multiples of it."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    blocks = -(-total // align)
    base, rem = divmod(blocks, world)
    lo_b = rank * base + min(rank, rem)
    hi_b = lo_b + base + (1 if rank < rem else 0)
    return min(lo_b * align, total), min(hi_b * align, total)


def allreduce_sum_(flat, group=None):
    """In-place SUM all-reduce of a flat gradient tensor; returns it.  The caller scales by 1/world
    (the reference's losses are batch means, gan_train_cwgangp_pixelnorm.py:215-216, so N ranks with
    batch B reproduce one rank with batch N*B)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat
