"""Multi-GPU plumbing: clips (and call streams) are independent, so the path shards with no
data-path collective; the only exchange is a gather of per-clip scores.
One process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from typing import List, Optional

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous, balanced shard [lo, hi) of n_items for `rank` (first n%world ranks get one more)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def stream_owner(stream_id: int, world: int) -> int:
    """Sticky stream -> rank map for live calls: a track's frame ring lives on one GPU."""
    return stream_id % world


def gather_scores(local: torch.Tensor, n_items: int, group=None) -> Optional[torch.Tensor]:
    """All-gather the ranks' score shards (made with shard_range) into the full [n_items]
    vector, in clip order, on every rank.  `local` is a 1-D float32 tensor on the backend's
    device (CUDA for NCCL, CPU for gloo)."""
    if not dist.is_available() or not dist.is_initialized():
        assert local.numel() == n_items
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(n_items, rank, world)
    assert local.numel() == hi - lo, (local.numel(), lo, hi)
    width = (n_items + world - 1) // world
    padded = torch.zeros(width, dtype=local.dtype, device=local.device)
    padded[: hi - lo] = local
    out: List[torch.Tensor] = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    parts = []
    for r in range(world):
        l, h = shard_range(n_items, r, world)
        parts.append(out[r][: h - l])
    return torch.cat(parts)


def max_over_ranks(value: float, device) -> float:
    if not dist.is_available() or not dist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
