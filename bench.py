#!/usr/bin/env python
"""bench.py -- PESQ + STOI/ESTOI audio-seconds scored per second on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--seconds S]

Workload (BASELINE.json configs[4]): PESQ + STOI/ESTOI of a batch of 8192 x 10 s synthetic
speech-like pairs at 16 kHz, batch-sharded over the N ranks (strong scaling: 8192 / N items per
GPU), only the per-item scores gathered with one NCCL all-gather.  One "step" = one PESQ call plus
one STOI call over the whole (sharded) batch = 81 920 audio-seconds.

  value  : device-resident inputs -> scores on device (+ all-gather), CUDA events on the launching
           stream, barrier + synchronize on both sides, max over ranks.
  e2e    : the same through the public API `PESQ(...)(clean, denoised)` / `STOI(...)(clean, denoised)`
           with pinned HOST tensors: the library's host entry point copies H->D in chunks overlapped
           with compute and copies the scores back, every step.
  roofline: dominant kernel (largest device time in the step, measured live with CUDA events around
           every launch): algorithmic bytes of its metric call (8 B per sample pair = 128 000 B per
           audio-second) / its average duration, against MEASURED_PEAKS.json's hbm_gbs.
  cpu_baseline / --impl reference: the numpy oracle port (oracle/) of the reference's CPU path on the
           host cores (one process per core), on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "PESQ+STOI audio-seconds scored/sec"
UNIT = "audio-s/s"
FS = 16000
BYTES_PER_AUDIO_SECOND_PER_METRIC = 2 * 4 * FS   # two fp32 signals read once


# ------------------------------------------------------------------------------------------ data
def make_shard(batch: int, n: int, seed: int, device):
    """Synthetic speech-like pairs (recipe of fast_speech_enhancement_metrics_b200.synth): a pool of
    seeded numpy items (low-passed noise x syllabic envelope + -60 dB floor), expanded on the GPU
    with a per-item circular shift and gain; degraded = clean + white noise at SNR ~ U[-5, 25] dB."""
    import numpy as np
    import torch

    from fast_speech_enhancement_metrics_b200.synth import synth_batch
    pool = min(batch, 256)
    base, _, _ = synth_batch(seed, pool, n)
    base = torch.from_numpy(base).to(device)
    g = torch.Generator(device=device).manual_seed(seed)
    clean = torch.empty(batch, n, dtype=torch.float32, device=device)
    for i0 in range(0, batch, pool):
        cnt = min(pool, batch - i0)
        shift = int(torch.randint(0, n, (1,), generator=g, device=device).item())
        gain = 0.5 + torch.rand(cnt, 1, generator=g, device=device)
        clean[i0:i0 + cnt] = torch.roll(base[:cnt], shifts=shift, dims=1) * gain
    deg = torch.empty_like(clean)
    snr = torch.rand(batch, 1, generator=g, device=device) * 30.0 - 5.0
    step = 512
    for i0 in range(0, batch, step):
        c = clean[i0:i0 + step]
        noise = torch.randn(c.shape, generator=g, device=device)
        scale = (c.pow(2).mean(1, keepdim=True) / 10.0 ** (snr[i0:i0 + step] / 10.0)).sqrt()
        deg[i0:i0 + step] = c + noise * scale
    return clean, deg


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU baseline
_CPU_DATA = None      # (clean[items, n], deg[items, n]) generated in the parent, inherited by the forked workers


_CPU_LIMITER = None


def _cpu_init():
    # one scoring thread per worker process: keep BLAS / OpenMP pools from oversubscribing the cores
    global _CPU_LIMITER
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        _CPU_LIMITER = threadpool_limits(limits=1)
    except Exception:
        pass
    from oracle import pesq_oracle, stoi_oracle  # noqa: F401  (pay the imports outside the timed region)


def _cpu_worker(i):
    from oracle import pesq_oracle, stoi_oracle
    clean, deg = _CPU_DATA
    p = pesq_oracle.pesq_batch(clean[i:i + 1], deg[i:i + 1])
    s, e, _ = stoi_oracle.stoi_batch(clean[i:i + 1], deg[i:i + 1], FS)
    return float(p[0]), float(s[0]), float(e[0])


class CpuOraclePool:
    """The numpy oracle port of the reference's CPU path (PESQ + STOI per item), one worker process per
    host core.  Inputs are generated once in the parent; the timed region is scoring only."""

    def __init__(self, n: int, items: int, cores: int):
        global _CPU_DATA
        import multiprocessing as mp

        from fast_speech_enhancement_metrics_b200.synth import synth_batch
        clean, deg, _ = synth_batch(4242, items, n)
        _CPU_DATA = (clean, deg)
        self.items, self.n, self.cores = items, n, cores
        self.pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init)
        self.pool.map(_cpu_worker, [i % items for i in range(2 * cores)], chunksize=1)      # warm every worker

    def step(self) -> float:
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, range(self.items), chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline(n: int, items: int, cores: int, repeats: int = 2) -> dict:
    pool = CpuOraclePool(n, items, cores)
    best = min(pool.step() for _ in range(repeats))
    pool.close()
    audio_s = items * n / FS
    return {"value": audio_s / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d items x %.0f s (PESQ+STOI per item, numpy oracle port of the reference CPU path, "
                      "%d worker processes, best of %d, inputs pre-generated)" % (items, n / FS, cores, repeats),
            "seconds": best}


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    n = int(args.seconds * FS)
    items = max(8 * cores, 64)
    pool = CpuOraclePool(n, items, cores)
    times = []
    for step in range(args.warmup + args.steps):
        dt = pool.step()
        if step >= args.warmup:
            times.append(dt)
    pool.close()
    audio_s = items * n / FS
    total = sum(times)
    value = audio_s * len(times) / total
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": {"workload": "PESQ+STOI/ESTOI, %d x %.0f s @16 kHz (BASELINE configs[4]); each step a bounded "
                               "sample of %d items" % (args.batch, args.seconds, items),
                   "sample_items_per_step": items},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d items x %.0f s per step, numpy oracle port of the reference CPU path "
                                   "(use_gpu=False), %d worker processes, inputs pre-generated"
                                   % (items, args.seconds, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------ main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8192, help="total items over all ranks")
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--two-streams", action="store_true",
                    help="experiment: run the PESQ and the STOI kernel chains on two CUDA streams")
    ap.add_argument("--overlap", type=int, default=-1,
                    help="experiment: fused device entry fsem_pesq_stoi_score_f32 with overlap mode 0/1/2/3 (3 = single-read first pass)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    # stdout carries exactly ONE JSON line (rank 0): anything libraries print there (e.g. NCCL's version banner) is
    # sent to stderr by pointing fd 1 at fd 2 until the line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line: dict):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)

    if args.impl == "reference":
        return run_reference(args, emit)

    import numpy as np
    import torch
    import torch.distributed as dist

    from fast_speech_enhancement_metrics_b200 import PESQ, STOI, _lib
    from fast_speech_enhancement_metrics_b200.dist import gather_scores, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    elif args.gpus > 1:
        print("bench.py: --gpus %d needs torchrun (WORLD_SIZE is 1); running 1 rank" % args.gpus, file=sys.stderr)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    n = int(args.seconds * FS)
    lo, hi = shard_range(args.batch, world, rank)
    local_b = hi - lo
    audio_s_total = args.batch * n / FS

    pesq = PESQ(FS, use_gpu=True)
    stoi = STOI(FS, use_gpu=True)
    clean, deg = make_shard(local_b, n, 1000 + rank, device)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gathered = torch.empty(args.batch, 3, dtype=torch.float32, device=device)

    side = (torch.cuda.Stream(device), torch.cuda.Stream(device)) if args.two_streams else None

    from fast_speech_enhancement_metrics_b200 import score_pesq_stoi_tensors

    def step_device():
        if args.overlap >= 0:
            sc3, _, _, _ = score_pesq_stoi_tensors(pesq, stoi, clean, deg, overlap=args.overlap)
            return gather_scores(sc3.t().contiguous(), args.batch, world, out=gathered)
        if side is None:
            mos, _ = pesq.score_tensors(clean, deg)
            sc, _, _ = stoi.score_tensors(clean, deg)
        else:
            cur = torch.cuda.current_stream(device)
            side[0].wait_stream(cur)
            side[1].wait_stream(cur)
            with torch.cuda.stream(side[0]):
                mos, _ = pesq.score_tensors(clean, deg)
            with torch.cuda.stream(side[1]):
                sc, _, _ = stoi.score_tensors(clean, deg)
            cur.wait_stream(side[0])
            cur.wait_stream(side[1])
        local = torch.stack([mos, sc[0], sc[1]], dim=1)
        return gather_scores(local, args.batch, world, out=gathered)

    # ---- value: device-resident inputs
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _lib.profile_reset()
    _lib.profile_enable(True)
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = step_device()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    _lib.profile_enable(False)
    prof = _lib.profile_read()
    if rank == 0 and len(sampler.lines) < 4:
        # a short timed region can end before nvidia-smi has sampled a few times: keep the same load running (untimed)
        # until there are enough samples to report the clocks under load
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end and len(sampler.lines) < 6:
            pesq.score_tensors(clean, deg)          # rank-local work only: no collective outside the timed steps
            stoi.score_tensors(clean, deg)
            torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = audio_s_total / (ms_per_step * 1e-3)
    scores = out.cpu().numpy()
    finite = float(np.isfinite(scores).mean())

    # ---- e2e: public API with pinned host tensors (H2D + compute + D2H every step)
    e2e = None
    if not args.no_e2e:
        hc = torch.empty(clean.shape, dtype=torch.float32, pin_memory=True).copy_(clean)
        hd = torch.empty(deg.shape, dtype=torch.float32, pin_memory=True).copy_(deg)
        torch.cuda.synchronize()

        from fast_speech_enhancement_metrics_b200 import score_pesq_stoi

        def finish(rows):
            local = torch.tensor(rows, dtype=torch.float32)
            if world > 1:
                return gather_scores(local.to(device), args.batch, world, out=gathered).cpu()
            return local

        def step_fused():        # one upload, both metrics (C ABI fsem_pesq_stoi_score_host_f32)
            r = score_pesq_stoi(pesq, stoi, hc, hd)
            return finish([[a["PESQ"], a["STOI"], a["ESTOI"]] for a in r])

        def step_separate():     # the reference's two calls, each uploading both signals
            p = pesq(hc, hd)
            s = stoi(hc, hd)
            return finish([[a["PESQ"], b["STOI"], b["ESTOI"]] for a, b in zip(p, s)])

        def time_e2e(fn, steps):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=device)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return audio_s_total * steps / float(dt.item())

        e_steps = max(1, min(args.steps, 3))
        v_fused = time_e2e(step_fused, e_steps)
        v_sep = time_e2e(step_separate, e_steps)
        del hc, hd
        # the same call on int16 PCM (ingest formats): 2 bytes per sample over PCIe, widened on
        # the device.  Informational -- the headline e2e above is the float32 contract of the reference API.
        scale = 32767.0 / float(torch.maximum(clean.abs().max(), deg.abs().max()))
        hc = torch.empty(clean.shape, dtype=torch.int16, pin_memory=True).copy_((clean * scale).round().to(torch.int16))
        hd = torch.empty(deg.shape, dtype=torch.int16, pin_memory=True).copy_((deg * scale).round().to(torch.int16))
        torch.cuda.synchronize()
        v_i16 = time_e2e(step_fused, e_steps)
        e2e = {"value": v_fused, "unit": UNIT,
               "h2d_bytes_per_step": int(2 * args.batch * n * 4),          # both signals, uploaded once
               "d2h_bytes_per_step": int(args.batch * 24),                 # mos, stoi, estoi, K, 2 x status
               "steps": e_steps,
               "api": "score_pesq_stoi(PESQ(16000), STOI(16000), clean_cpu, deg_cpu) with pinned host tensors "
                      "(C ABI fsem_pesq_stoi_score_host_f32: one upload, both metrics)",
               "separate_calls": {"value": v_sep, "unit": UNIT, "h2d_bytes_per_step": int(2 * 2 * args.batch * n * 4),
                                  "api": "PESQ(16000)(clean_cpu, deg_cpu) + STOI(16000)(clean_cpu, deg_cpu): "
                                         "the reference's two calls, each uploading both signals"},
               "int16_ingest": {"value": v_i16, "unit": UNIT, "h2d_bytes_per_step": int(2 * args.batch * n * 2),
                                "api": "score_pesq_stoi on pinned int16 PCM tensors (C ABI fsem_score_host, "
                                       "FSEM_DTYPE_I16): same samples quantised to 16 bit, widened on the device"}}
        del hc, hd

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    kernels = {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps} for k, v in prof.items()}
    top = max(prof.items(), key=lambda kv: kv[1][0])
    top_name, (top_ms, top_cnt) = top
    local_audio_s = local_b * n / FS
    alg_bytes = BYTES_PER_AUDIO_SECOND_PER_METRIC * local_audio_s          # per launch (one launch per metric call)
    top_avg_ms = top_ms / max(top_cnt, 1)
    achieved = alg_bytes / (top_avg_ms * 1e-3) / 1e9
    traffic_path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    traffic = None
    if os.path.exists(traffic_path):
        tj = json.load(open(traffic_path))
        per_launch = tj.get("bytes_per_launch", {}).get(top_name)
        if per_launch is not None:          # measured with ncu at tj["items"] items x 10 s; linear in the item count
            traffic = per_launch * (local_b * args.seconds) / (tj.get("items", 8192) * 10.0)
    roofline = {"bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": top_avg_ms,
                "note": "FP32-issue-bound path (SURVEY 7.2): the HBM fraction is reported as the metric asks"}
    step_bytes = 2 * BYTES_PER_AUDIO_SECOND_PER_METRIC * local_audio_s
    roofline_step = {"achieved": step_bytes / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
                     "algorithmic_bytes_per_step_per_gpu": step_bytes}

    cpu = None
    if not args.no_cpu and world == 1:
        cores = host_cores()
        cpu = cpu_baseline(n, max(64, 32 * cores), cores)      # ~20-30 CPU-seconds of scoring

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "PESQ + STOI/ESTOI, %d x %.0f s @16 kHz, batch-sharded over %d GPU(s) "
                               "(BASELINE configs[4])" % (args.batch, args.seconds, world),
                   "batch_total": args.batch, "batch_per_gpu": local_b, "samples": n, "sample_rate": FS,
                   "l2": "inputs (%.1f GB per GPU) exceed L2; no flush" % (2 * local_b * n * 4 / 1e9),
                   "collective": "all_gather of [batch/N, 3] fp32 scores (NCCL)" if world > 1 else "none"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "roofline_step": roofline_step, "kernels": kernels, "cpu_baseline": cpu,
        "finite_score_fraction": finite,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
