"""Pair sharding for multi-GPU inference: one process per GPU, contiguous ranges of image pairs, no data-path
collective (SURVEY.md §8e — pairs are independent in eval mode).  The only communication is the optional
max-reduce of timings / gather of small host-side results."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_pairs(total_pairs: int, rank: int, world: int):
    """Contiguous [lo, hi) range of pairs for `rank`; the first `total % world` ranks take one extra pair."""
    base, rem = divmod(total_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_max_ms(ms: float, device="cuda") -> float:
    """max over ranks of a duration (the step time of a sharded job is its slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms)
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def infer_sharded(net, x1_all, x2_all):
    """Strong-scaling helper: every rank holds (or can index) the global batch on the host; it uploads and runs
    only its own contiguous shard and returns (lo, hi, logits_of_shard).  No collective."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_pairs(x1_all.shape[0], rank, world)
    dev = next(net.parameters()).device
    with torch.no_grad():
        y = net(x1_all[lo:hi].to(dev, non_blocking=True), x2_all[lo:hi].to(dev, non_blocking=True))
    return lo, hi, y
