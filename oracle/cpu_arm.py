"""CPU arm of bench.py: the reference's CPU path (or the numpy port) timed on the host cores.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by bench.py's `cpu_baseline` leg and by
`bench.py --impl reference`, never by the product.

Two kinds:
  "reference"  oracle/_ref = the UNMODIFIED reference modules (oracle/make_ref.py), called through their public API
               exactly as BASELINE.md section 3 prescribes: `PESQ(16000, use_gpu=False)(clean, deg)` and
               `STOI(16000, use_gpu=False)(clean, deg)` on 64-item chunks of float32 CPU tensors.
  "port"       the float64 numpy restatement (oracle/pesq_oracle.py, stoi_oracle.py), per item.

The reference parallelises only inside torch ops (one process, `torch.get_num_threads()` intra-op threads), which
leaves most cores idle in its Python glue; "all the host threads it can use" is therefore taken literally: `procs`
worker processes (spawned, so no CUDA / OpenMP state is inherited), each with `threads` intra-op threads, each
scoring its own chunk per step.  The stock single-process configuration is timed beside it (`single_process`).
Inputs are generated per worker before timing (same synthetic recipe as the GPU arm, seeded per worker); the
timed region of a step is scoring only.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FS = 16000
_STATE = {}


def _worker_init(kind: str, n: int, items: int, threads: int, seed: int):
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ["MKL_NUM_THREADS"] = str(threads)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    ident = mp.current_process()._identity
    wid = ident[0] if ident else 0
    from fast_speech_enhancement_metrics_b200.synth import synth_batch
    clean, deg, _ = synth_batch(seed + wid, items, n)
    _STATE.update(kind=kind, clean=clean, deg=deg)
    if kind == "reference":
        import torch
        torch.set_num_threads(threads)
        from oracle import make_ref
        PESQ, STOI = make_ref.load()
        _STATE.update(pesq=PESQ(FS, use_gpu=False), stoi=STOI(FS, use_gpu=False),
                      tclean=torch.from_numpy(clean), tdeg=torch.from_numpy(deg), threads=torch.get_num_threads())
    else:
        try:
            from threadpoolctl import threadpool_limits
            _STATE["limiter"] = threadpool_limits(limits=threads)
        except Exception:
            pass
        from oracle import pesq_oracle, stoi_oracle  # noqa: F401
    _score(min(2, items))                               # warm: imports, filter design, thread pools


def _score(count=None):
    """PESQ + STOI/ESTOI of this worker's chunk; returns [(pesq, stoi, estoi), ...]."""
    kind = _STATE["kind"]
    if kind == "reference":
        c, d = _STATE["tclean"], _STATE["tdeg"]
        if count is not None:
            c, d = c[:count], d[:count]
        p = _STATE["pesq"](c, d)
        s = _STATE["stoi"](c, d)
        return [(a["PESQ"], b["STOI"], b["ESTOI"]) for a, b in zip(p, s)]
    from oracle import pesq_oracle, stoi_oracle
    c, d = _STATE["clean"], _STATE["deg"]
    if count is not None:
        c, d = c[:count], d[:count]
    p = pesq_oracle.pesq_batch(c, d)
    s, e, _ = stoi_oracle.stoi_batch(c, d, FS)
    return list(zip(p.tolist(), s.tolist(), e.tolist()))


def _worker_step(_):
    t0 = time.perf_counter()
    out = _score()
    return time.perf_counter() - t0, len(out), sum(1 for r in out if all(v == v for v in r))


class CpuArm:
    """`procs` workers x `threads` intra-op threads; one step = every worker scores its own `items`-item chunk."""

    def __init__(self, kind: str, n: int, procs: int, threads: int, items: int, seed: int = 4242):
        self.kind, self.n, self.procs, self.threads, self.items = kind, n, procs, threads, items
        self.pool = mp.get_context("spawn").Pool(procs, initializer=_worker_init, initargs=(kind, n, items, threads, seed))
        self.pool.map(_worker_step, range(procs), chunksize=1)          # every worker up and warm before timing

    def step(self) -> float:
        t0 = time.perf_counter()
        res = self.pool.map(_worker_step, range(self.procs), chunksize=1)
        dt = time.perf_counter() - t0
        self.finite = sum(r[2] for r in res) / max(1, sum(r[1] for r in res))
        return dt

    @property
    def audio_seconds_per_step(self) -> float:
        return self.procs * self.items * self.n / FS

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def pick_kind() -> str:
    from oracle import make_ref
    return "reference" if make_ref.available() else "port"


def plan(kind: str, cores: int):
    """(procs, threads, items per worker and step).  The reference runs 64-item chunks (BASELINE.md section 3;
    STOI materialises [2B, M, 15, 30]: ~1.5 GB live per 64-item chunk, so the worker count is also capped by RAM);
    the port scores item by item, 8 items per worker and step."""
    if kind == "reference":
        procs = cores
        try:
            import psutil
            procs = max(1, min(procs, int(psutil.virtual_memory().available * 0.6 / 2.5e9)))
        except Exception:
            pass
        return procs, 1, 64
    return cores, 1, 8


def measure(kind: str, n: int, cores: int, repeats: int = 3, single_process: bool = True) -> dict:
    """cpu_baseline object of the bench line: best of `repeats` steps after a warm-up step."""
    procs, threads, items = plan(kind, cores)
    arm = CpuArm(kind, n, procs, threads, items)
    best = min(arm.step() for _ in range(repeats))
    finite = arm.finite
    arm.close()
    out = {"value": arm.audio_seconds_per_step / best, "unit": "audio-s/s", "cores": cores, "kind": kind,
           "procs": procs, "threads_per_proc": threads, "seconds": best, "finite_score_fraction": finite,
           "sample": "%d worker processes x %d items x %.0f s per step (PESQ+STOI/ESTOI), %s, best of %d after warm-up, "
                     "inputs pre-generated" % (procs, items, n / FS, _describe(kind, threads), repeats)}
    if single_process and kind == "reference":
        arm = CpuArm(kind, n, 1, cores, items)           # the stock configuration: one process, torch intra-op threads
        b1 = min(arm.step() for _ in range(repeats))
        arm.close()
        out["single_process"] = {"value": arm.audio_seconds_per_step / b1, "unit": "audio-s/s", "procs": 1,
                                 "torch_num_threads": cores, "seconds": b1,
                                 "sample": "one process, torch.set_num_threads(%d), one %d-item chunk" % (cores, items)}
    return out


def _describe(kind: str, threads: int) -> str:
    if kind == "reference":
        return ("the UNMODIFIED reference (oracle/_ref): PESQ(16000, use_gpu=False) + STOI(16000, use_gpu=False) on "
                "64-item chunks, torch.set_num_threads(%d) per process" % threads)
    return "numpy float64 port of the reference CPU path (oracle/), per item"
