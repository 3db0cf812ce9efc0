#!/usr/bin/env python
"""Device-resident timing of every BASELINE.json config (and a batch-size ladder) on one GPU.

    python tools/config_sweep.py [--out gpurun_out/config_sweep.json]

For each shape: PESQ and STOI/ESTOI calls on CUDA tensors, CUDA events over `steps` calls after warm-up, plus the
per-kernel times from the library's own profiler.  bench.py stays the single bench line the driver reads; this
script only documents how the path behaves away from the headline shape (small batches are latency-bound).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_speech_enhancement_metrics_b200 import LSD, PESQ, SDR, STOI, CapturedScorer, _lib  # noqa: E402
from fast_speech_enhancement_metrics_b200.synth import synth_batch  # noqa: E402


def make(batch, n, seed):
    base = min(batch, 64)
    c, d, _ = synth_batch(seed, base, n)
    c, d = torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda()
    if batch > base:                                  # tile with per-row circular shifts: distinct rows, same statistics
        reps = (batch + base - 1) // base
        c = torch.cat([torch.roll(c, 977 * r, dims=1) for r in range(reps)])[:batch].contiguous()
        d = torch.cat([torch.roll(d, 977 * r, dims=1) for r in range(reps)])[:batch].contiguous()
    return c, d


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    _lib.profile_reset()
    _lib.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    _lib.profile_enable(False)
    prof = {k: round(v[0] / steps, 4) for k, v in _lib.profile_read().items() if v[1]}
    return e0.elapsed_time(e1) / steps, prof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/config_sweep.json")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--max-batch", type=int, default=0, help="only shapes up to this batch size (0 = all); skips LSD / SDR")
    args = ap.parse_args()
    pesq, stoi = PESQ(16000, use_gpu=True), STOI(16000, use_gpu=True)
    shapes = [("configs[0] README 4 x 10 s", 4, 160000, None),
              ("configs[1] PESQ 256 x 10 s", 256, 160000, None),
              ("configs[2] STOI 1024 x 4 s", 1024, 64000, None),
              ("configs[3] variable 1-30 s, 1024 items", 1024, 480000, "var"),
              ("target statement 4096 x 10 s", 4096, 160000, None),
              ("configs[4] 8192 x 10 s", 8192, 160000, None)]
    shapes += [("ladder %d x 10 s" % b, b, 160000, None) for b in (1, 2, 8, 16, 32, 64, 128, 512, 1024, 2048)]
    rows = []
    for i, (name, b, n, mode) in enumerate(shapes):
        if args.max_batch and b > args.max_batch:
            continue
        c, d = make(b, n, 2000 + i)
        lengths = None
        if mode == "var":
            lengths = torch.from_numpy(np.random.default_rng(7).integers(16000, n + 1, size=b).astype(np.int32)).cuda()
        audio_s = float(lengths.sum()) / 16000.0 if lengths is not None else b * n / 16000.0
        row = {"shape": name, "batch": b, "samples": n, "audio_s": audio_s}
        for label, metric in (("PESQ", pesq), ("STOI", stoi)):
            ms, prof = timed(lambda: metric.score_tensors(c, d, lengths), args.steps)
            row[label] = {"ms": round(ms, 4), "audio_s_per_s": round(audio_s / (ms * 1e-3), 1), "kernels_ms": prof}
        both = row["PESQ"]["ms"] + row["STOI"]["ms"]
        row["PESQ+STOI"] = {"ms": round(both, 4), "audio_s_per_s": round(audio_s / (both * 1e-3), 1)}
        # the same scoring as ONE CUDA graph (parallel branches; large batches also cut into slices): CUDA events
        # around `steps` replays
        def graph_ms(slices):
            scorer = CapturedScorer(pesq, stoi, c, d, lengths, slices=slices)
            for _ in range(3):
                scorer.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = args.steps * (20 if b <= 64 else 1)
            e0.record()
            for _ in range(reps):
                scorer.replay()
            e1.record()
            torch.cuda.synchronize()
            out = (e0.elapsed_time(e1) / reps, scorer.kernel_nodes, scorer.slices)
            scorer.close()
            return out
        gms, nodes, used = graph_ms(0)
        row["PESQ+STOI graph"] = {"ms": round(gms, 4), "audio_s_per_s": round(audio_s / (gms * 1e-3), 1), "slices": used,
                                  "kernel_nodes": nodes, "speedup_vs_calls": round(both / gms, 3)}
        if b >= 512:
            row["PESQ+STOI graph"]["by_slices"] = {str(k): round(graph_ms(k)[0], 4) for k in (1, 2, 3, 4, 8)}
            print("   by slices:", row["PESQ+STOI graph"]["by_slices"], flush=True)
        rows.append(row)
        print("%-42s PESQ %9.3f ms  STOI %9.3f ms  both %12.0f audio-s/s   graph %9.3f ms (x%.2f)" % (
            name, row["PESQ"]["ms"], row["STOI"]["ms"], row["PESQ+STOI"]["audio_s_per_s"], gms, both / gms), flush=True)
        del c, d
        torch.cuda.empty_cache()
    # adjacent metrics at the headline shape, with their HBM roofline line: one call must read both
    # fp32 signals once = 128 000 B per audio-second
    peak = 6555.8
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                 "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    lsd, sdr = LSD(16000, use_gpu=True), SDR(16000, use_gpu=True)
    for name, b, n in (("LSD/SDR 8192 x 10 s", 8192, 160000), ("LSD/SDR 256 x 10 s", 256, 160000)):
        if args.max_batch:
            break
        c, d = make(b, n, 3000 + b)
        audio_s = b * n / 16000.0
        row = {"shape": name, "batch": b, "samples": n, "audio_s": audio_s}
        for label, metric in (("LSD", lsd), ("SDR", sdr)):
            ms, prof = timed(lambda: metric.score_tensors(c, d), max(2, args.steps // 3))
            gbs = 128000.0 * audio_s / (ms * 1e-3) / 1e9
            row[label] = {"ms": round(ms, 4), "audio_s_per_s": round(audio_s / (ms * 1e-3), 1), "kernels_ms": prof,
                          "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s",
                                       "frac": round(gbs / peak, 4)}}
        rows.append(row)
        print("%-42s LSD  %9.3f ms  SDR  %9.3f ms" % (name, row["LSD"]["ms"], row["SDR"]["ms"]), flush=True)
        del c, d
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
