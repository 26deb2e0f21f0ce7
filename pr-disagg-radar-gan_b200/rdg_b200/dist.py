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


def broadcast_(t, src=0, group=None):
    """In-place broadcast of a tensor from rank `src` (replica initialisation of data-parallel training)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(t, src=src, group=group)
    return t


def bind_to_gpu_numa_node(device_index: int):
    """Pin the calling process to the CPUs of the NUMA node its GPU hangs off (Linux; best effort, returns the node or None).
    The host-buffer paths move 24 KB per scenario over PCIe into pinned host memory; with one process per GPU on a two-socket
    box an unpinned process lands its staging buffers on the remote socket half of the time and the aggregate D2H rate of 8
    ranks drops to about half.  Call before allocating pinned memory (first touch then places it on the local node)."""
    import os
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0"
        with open(os.path.join(path, "numa_node")) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None
