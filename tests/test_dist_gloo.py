"""World-size-2 gloo tests (CPU) of the sharding / score-gather host logic."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fast_speech_enhancement_metrics_b200.dist import (balanced_order, deal_round_robin, gather_scores,
                                                       shard_range, shard_ranges)


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
