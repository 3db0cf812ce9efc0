"""NCCL test (-m gpu, needs >= 2 GPUs; skipped otherwise): the batch-sharded scoring path of bench.py / dist.py over
real NCCL with an UNEVEN batch (8191 items over 2 ranks: shards of 4096 and 4095, padded for the collective).

Every rank scores its own shard with the CUDA kernels, the score rows are all-gathered, and every rank checks the
gathered [8191, 3] table bit for bit against its own scoring of BOTH shards (identical GPUs, deterministic kernels),
plus a sample of rows against the float64 oracle."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BATCH = 8191
N = 24000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_batch(device):
    from fast_speech_enhancement_metrics_b200.synth import synth_batch
    clean, deg, _ = synth_batch(909, 32, N)
    reps = -(-BATCH // 32)
    c = torch.from_numpy(clean).to(device).repeat(reps, 1)[:BATCH]
    d = torch.from_numpy(deg).to(device).repeat(reps, 1)[:BATCH]
    gains = torch.linspace(0.25, 2.0, BATCH, device=device)[:, None]
    return (c * gains).contiguous(), (d * gains).contiguous(), clean, deg


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist

        from fast_speech_enhancement_metrics_b200 import PESQ, STOI
        from fast_speech_enhancement_metrics_b200.dist import gather_scores, shard_ranges
        from oracle import pesq_oracle, stoi_oracle
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        device = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
        pesq, stoi = PESQ(16000, use_gpu=True), STOI(16000, use_gpu=True)
        c, d, _, _ = _make_batch(device)
        ranges = shard_ranges(BATCH, world)

        def score(lo, hi):
            mos, _ = pesq.score_tensors(c[lo:hi], d[lo:hi])
            sc, _, _ = stoi.score_tensors(c[lo:hi], d[lo:hi])
            return torch.stack([mos, sc[0], sc[1]], dim=1)

        lo, hi = ranges[rank]
        full = gather_scores(score(lo, hi), BATCH, world)
        want = torch.cat([score(a, b) for a, b in ranges])
        ok_shape = tuple(full.shape) == (BATCH, 3)
        ok_bits = bool(torch.equal(full.contiguous().view(torch.int32), want.contiguous().view(torch.int32)))
        pick = [0, 31, 4095, 4096, BATCH - 1]
        cs, ds = c[pick].cpu().numpy(), d[pick].cpu().numpy()
        got = full[pick].double().cpu().numpy()
        def maxabs(a, b):      # NaN (an item without a 30-frame segment) must be NaN on both sides
            assert np.array_equal(np.isnan(a), np.isnan(b)), (a, b)
            return float(np.nanmax(np.abs(a - b))) if not np.all(np.isnan(a)) else 0.0

        dp = maxabs(got[:, 0], pesq_oracle.pesq_batch(cs, ds))
        ws, we, _ = stoi_oracle.stoi_batch(cs, ds, 16000)
        dsd = max(maxabs(got[:, 1], ws), maxabs(got[:, 2], we))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, ok_shape, ok_bits, dp, dsd))
    except Exception as exc:   # report instead of hanging the parent
        q.put((rank, False, False, repr(exc), 0.0))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL)")
def test_sharded_scores_over_nccl_uneven_batch():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    for rank, ok_shape, ok_bits, dp, dsd in res:
        assert ok_shape and ok_bits, (rank, dp)
        assert dp <= 2e-4 and dsd <= 1e-4, (rank, dp, dsd)


def _ragged_worker(rank, world, port, q):
    try:
        import torch.distributed as dist

        from fast_speech_enhancement_metrics_b200 import PESQ, STOI, score_pesq_stoi_tensors
        from fast_speech_enhancement_metrics_b200.dist import dealt_parts, score_sharded
        from fast_speech_enhancement_metrics_b200.synth import synth_batch
        from oracle import pesq_oracle, stoi_oracle
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        device = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
        pesq, stoi = PESQ(16000, use_gpu=True), STOI(16000, use_gpu=True)
        b, n = 203, 48000
        clean, deg, _ = synth_batch(4711, b, n)                     # the same host batch on every rank
        lens = np.random.default_rng(5).integers(8000, n + 1, size=b)
        lens[3], lens[-1] = n, 8000
        lens = lens.tolist()
        c, d = torch.from_numpy(clean), torch.from_numpy(deg)
        full = score_sharded(pesq, stoi, c, d, lens)
        # every rank re-scores each rank's dealt part with the same local batch -> identical kernels and chunk grids
        parts = dealt_parts(b, world, lens)
        want = torch.empty(b, 3, dtype=torch.float32, device=device)
        for p in parts:
            idx = torch.as_tensor(p)
            s, _, _, _ = score_pesq_stoi_tensors(pesq, stoi, c[idx].to(device), d[idx].to(device), [lens[i] for i in p])
            want[idx.to(device)] = s.t()
        ok_bits = bool(torch.equal(full.contiguous().view(torch.int32), want.contiguous().view(torch.int32)))
        sums = [sum(lens[i] for i in p) for p in parts]
        balanced = (max(sums) - min(sums)) / max(sums) < 0.05
        pick = [0, 3, 100, b - 1]
        got = full[pick].double().cpu().numpy()
        pl = [lens[i] for i in pick]
        dp = float(np.max(np.abs(got[:, 0] - pesq_oracle.pesq_batch(clean[pick], deg[pick], pl))))
        ws, we, _ = stoi_oracle.stoi_batch(clean[pick], deg[pick], 16000, pl)
        dsd = float(max(np.nanmax(np.abs(got[:, 1] - ws)), np.nanmax(np.abs(got[:, 2] - we))))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, tuple(full.shape) == (b, 3) and balanced, ok_bits, dp, dsd))
    except Exception as exc:   # report instead of hanging the parent
        q.put((rank, False, False, repr(exc), 0.0))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL)")
def test_ragged_batch_dealt_by_length_over_nccl():
    """Variable-length batch: items sorted by length and dealt round-robin over the ranks, scored with the
    fused entry point, rows gathered back into the caller's order (dist.score_sharded)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ragged_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    for rank, ok_shape, ok_bits, dp, dsd in res:
        assert ok_shape and ok_bits, (rank, dp)
        assert dp <= 2e-4 and dsd <= 1e-4, (rank, dp, dsd)
