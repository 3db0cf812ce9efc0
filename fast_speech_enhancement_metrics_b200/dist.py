"""Batch sharding across the GPUs of one box.

Utterances are independent (the reference has no cross-item op: SURVEY.md 8e), so the
batch is cut into contiguous slices, one per rank, every rank scores its slice with the
same kernels, and the only exchange is one all-gather of the per-item score rows
(NCCL over NVLink on GPUs; gloo in the CPU tests).  Uneven shards are padded to the
largest shard for the collective and trimmed afterwards.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [lo, hi) slice of rank `rank`: ceil(batch / world) items per rank
    (the last ranks may get fewer, possibly none)."""
    per = -(-batch // world)
    lo = min(rank * per, batch)
    return lo, min(lo + per, batch)


def shard_ranges(batch: int, world: int) -> list[tuple[int, int]]:
    return [shard_range(batch, world, r) for r in range(world)]


def balanced_order(lengths) -> list[int]:
    """Item order for variable-length batches: sort by length (longest first) and deal
    round-robin, so that contiguous shards of the reordered batch carry near-equal
    sample counts.  Returns the permutation (new position -> original index)."""
    idx = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    return idx


def deal_round_robin(order: list[int], world: int) -> list[list[int]]:
    """Rank r scores items order[r], order[r + world], ... (longest-first dealing)."""
    return [order[r::world] for r in range(world)]


def gather_scores(local: torch.Tensor, batch: int, world: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """All-gather the [local_batch, C] score rows of contiguous shards into [batch, C].
    With world == 1 this is a copy-free pass-through."""
    if world == 1:
        return local
    per = -(-batch // world)
    c = local.shape[1]
    if local.shape[0] == per and batch == per * world:
        if out is None:
            out = torch.empty(batch, c, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    padded = torch.zeros(per, c, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty(per * world, c, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded)
    return buf[:batch]
