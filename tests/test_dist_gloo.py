"""World-size-2 gloo tests (CPU) of the sharding / score-gather host logic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fast_speech_enhancement_metrics_b200.dist import (balanced_order, deal_round_robin, dealt_parts, gather_dealt,
                                                       gather_scores, shard_range, shard_ranges)


def test_shard_ranges_cover_the_batch():
    for batch in (1, 7, 8, 8192, 8193):
        for world in (1, 2, 3, 4, 8):
            r = shard_ranges(batch, world)
            assert r[0][0] == 0 and r[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(hi - lo for lo, hi in r) == -(-batch // world)


def test_balanced_dealing():
    lengths = [16000 * (1 + (i * 7) % 30) for i in range(64)]
    order = balanced_order(lengths)
    parts = deal_round_robin(order, 4)
    assert sorted(sum(parts, [])) == list(range(64))
    sums = [sum(lengths[i] for i in p) for p in parts]
    assert (max(sums) - min(sums)) / max(sums) < 0.12


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(batch, world, rank)
    # the "scores" of item i are (i, 2i, 3i): every rank must end with the full table in order
    idx = torch.arange(lo, hi, dtype=torch.float32)
    local = torch.stack([idx, 2 * idx, 3 * idx], dim=1)
    full = gather_scores(local, batch, world)
    want = torch.arange(batch, dtype=torch.float32)
    ok = full.shape == (batch, 3) and torch.equal(full[:, 0], want) and torch.equal(full[:, 2], 3 * want)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_gather_scores_world2_gloo(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_dealt_parts_balance_and_cover():
    lengths = [16000 * (1 + (i * 7) % 30) for i in range(101)]
    for world in (1, 2, 3, 8):
        parts = dealt_parts(101, world, lengths)
        assert sorted(sum(parts, [])) == list(range(101))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        sums = [sum(lengths[i] for i in p) for p in parts]
        assert (max(sums) - min(sums)) / max(sums) < 0.15
        assert dealt_parts(101, world) == [list(range(lo, hi)) for lo, hi in shard_ranges(101, world)]
    # one process, no collective: the rows come back in the caller's order
    parts = dealt_parts(7, 1, [5, 3, 9, 1, 7, 2, 8])
    local = torch.tensor([[float(i), 10.0 * i] for i in parts[0]])
    full = gather_dealt(local, parts, 0)
    assert torch.equal(full[:, 0], torch.arange(7, dtype=torch.float32))


def _dealt_worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lengths = [1000 + 37 * ((i * 11) % 23) for i in range(batch)]
    parts = dealt_parts(batch, world, lengths)
    idx = torch.tensor(parts[rank], dtype=torch.float32)
    local = torch.stack([idx, 2 * idx, 3 * idx], dim=1)          # the "scores" of item i are (i, 2i, 3i)
    full = gather_dealt(local, parts, rank)
    want = torch.arange(batch, dtype=torch.float32)
    ok = full.shape == (batch, 3) and torch.equal(full[:, 0], want) and torch.equal(full[:, 2], 3 * want)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [9, 16])
def test_gather_dealt_world2_gloo(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dealt_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
