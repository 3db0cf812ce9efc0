"""Batch sharding across the GPUs of one box.

Utterances are independent (the reference has no cross-item op: SURVEY.md 8e), so the
batch is cut into contiguous slices, one per rank, every rank scores its slice with the
same kernels, and the only exchange is one all-gather of the per-item score rows
(NCCL over NVLink on GPUs; gloo in the CPU tests).  Uneven shards are padded to the
largest shard for the collective and trimmed afterwards.

Variable-length batches (per-item `lengths`): contiguous slices would give the ranks unequal
work, so the items are sorted by length and dealt round-robin (`dealt_parts`); every rank
scores its own items, the score rows are all-gathered and put back in the caller's order
(`gather_dealt`).  `score_sharded` is the whole call for a batch every rank can see.
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


def dealt_parts(batch: int, world: int, lengths=None) -> list[list[int]]:
    """Item indices per rank.  Without lengths: the contiguous slices of shard_ranges.  With lengths: longest-first
    round-robin dealing, so every rank gets the same item count (+-1) and near-equal total samples."""
    if lengths is None:
        return [list(range(lo, hi)) for lo, hi in shard_ranges(batch, world)]
    if len(lengths) != batch:
        raise Exception("`lengths` must have one entry per batch item")
    return deal_round_robin(balanced_order(lengths), world)


def gather_dealt(local: torch.Tensor, parts: list[list[int]], rank: int) -> torch.Tensor:
    """All-gather the [len(parts[rank]), C] score rows of every rank and return the [batch, C] table in the ORIGINAL item
    order on every rank.  One all_gather_into_tensor of equal (padded) blocks, then one index scatter."""
    world = len(parts)
    batch = sum(len(p) for p in parts)
    if local.shape[0] != len(parts[rank]):
        raise Exception("rank %d scored %d rows, its part has %d items" % (rank, local.shape[0], len(parts[rank])))
    c = local.shape[1]
    if world == 1:
        out = torch.empty(batch, c, dtype=local.dtype, device=local.device)
        out[torch.as_tensor(parts[0], dtype=torch.long, device=local.device)] = local
        return out
    per = max(len(p) for p in parts)
    padded = torch.zeros(per, c, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty(per * world, c, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded)
    src = torch.as_tensor([r * per + i for r, p in enumerate(parts) for i in range(len(p))], dtype=torch.long,
                          device=local.device)
    dst = torch.as_tensor([i for p in parts for i in p], dtype=torch.long, device=local.device)
    out = torch.empty(batch, c, dtype=local.dtype, device=local.device)
    out[dst] = buf[src]
    return out


def score_sharded(pesq, stoi, clean: torch.Tensor, deg: torch.Tensor, lengths=None) -> torch.Tensor:
    """PESQ / STOI / ESTOI of a batch that every rank holds (host or device tensors), scored once across the ranks of
    the default process group: rank r scores its dealt items with the fused entry point, the rows are gathered.  Returns
    the [batch, 3] table (columns PESQ, STOI, ESTOI) on the metric's device, identical on every rank and -- the kernels
    being deterministic and per-item -- identical to scoring the whole batch on one GPU, except for PESQ's IIR chunk
    grid, which depends on the local batch size (differences ~1e-7 in the filtered signal)."""
    from .fused import score_pesq_stoi_tensors
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    batch = clean.shape[0]
    lens = None if lengths is None else [int(x) for x in torch.as_tensor(lengths).reshape(-1).tolist()]
    parts = dealt_parts(batch, world, lens)
    mine = parts[rank]
    if len(mine) == 0:
        local = torch.empty(0, 3, dtype=torch.float32, device=pesq.device)
    else:
        if lens is None:                                   # contiguous slice: a view, no copy
            c, d = clean[mine[0]:mine[-1] + 1], deg[mine[0]:mine[-1] + 1]
        else:
            idx = torch.as_tensor(mine, dtype=torch.long, device=clean.device)
            c, d = clean.index_select(0, idx), deg.index_select(0, idx)
        c, d = c.to(pesq.device, non_blocking=True), d.to(pesq.device, non_blocking=True)
        scores, _, _, _ = score_pesq_stoi_tensors(pesq, stoi, c, d, None if lens is None else [lens[i] for i in mine])
        local = scores.t().contiguous()
    return gather_dealt(local, parts, rank)
