"""Multi-GPU plumbing for one seed / expansion wave: the frontier is cut into contiguous
shards, one per rank (images and cameras are replicated); after the refine kernel every rank contributes its
refined patch records to one all-gather so that all ranks apply the same deterministic cell-occupancy commit.
No collective touches the texel data path."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced (sizes differ by at most one) partition of range(n)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allgather_records(rec: torch.Tensor, n_total: int, world: int) -> torch.Tensor:
    """rec: (n_local, K) records of this rank's shard, in shard order.  Returns the (n_total, K) table in global
    order on every rank.  Shards may differ by one row, so rows are padded to the largest shard for the
    fixed-size collective and trimmed afterwards."""
    if world == 1:
        return rec
    rank = dist.get_rank()
    k = rec.shape[1]
    mx = (n_total + world - 1) // world
    buf = torch.zeros(mx, k, dtype=rec.dtype, device=rec.device)
    buf[: rec.shape[0]] = rec
    out = torch.empty(world * mx, k, dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, buf)
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        parts.append(out[r * mx: r * mx + (hi - lo)])
    assert shard_bounds(n_total, rank, world)[1] - shard_bounds(n_total, rank, world)[0] == rec.shape[0]
    return torch.cat(parts, dim=0)
