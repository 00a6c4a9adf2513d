"""Sharding a query batch across the GPUs of one box.

Every query point is independent (`Interpolator::interpolate(&self, ..)`, three/mod.rs:193), so the batch is
split into contiguous index ranges, one per rank, with the grid replicated on every GPU and no collective on the
data path.  The optional gather of result shards uses `torch.distributed` (NCCL on GPUs, gloo in the CPU tests)
and is not part of the timed path.
"""
from __future__ import annotations

from typing import Sequence


def shard_range(n: int, rank: int, world_size: int) -> tuple[int, int]:
    """[lo, hi) of rank's contiguous share of n points; shares differ by at most one point."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} for world size {world_size}")
    base, extra = divmod(int(n), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n: int, world_size: int) -> list[int]:
    return [shard_range(n, r, world_size)[1] - shard_range(n, r, world_size)[0] for r in range(world_size)]


def shard_columns(cols: Sequence, rank: int, world_size: int):
    """Rank's slice of every coordinate column (tensors or arrays)."""
    lo, hi = shard_range(len(cols[0]), rank, world_size)
    return [c[lo:hi] for c in cols]


def gather_results(local, n_total: int, group=None):
    """all_gather of ragged result shards (values or status bytes) into one tensor of n_total elements.

    `local` is this rank's shard as produced by `interpolate_batch` on `shard_columns(...)`.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = shard_sizes(n_total, world)
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=local.dtype, device=local.device)
    buf[: local.numel()] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])
