#!/usr/bin/env python
"""Per-batch-size timing harness writing the reference's results schema.

The reference's `benchmark_metrics.py` (its :81-108) stores, for every batch size,
`results/batch_size_<B>/<Metric>_results.json` with one entry per implementation
(`"<Class>_GPU"`: `{"batch_times": [...], "values": [...]}`) plus `snrs`, `batch_size`,
`sample_duration`, `sample_rate`, `SNR_high`, `SNR_low`; its plotting scripts
(`benchmarking/plotting/utils.py:8-35`) read exactly these keys.  This harness produces the same
files for the two metrics of this repository so those scripts work unchanged.  Differences from the
reference driver: synthetic speech-like audio instead of the streamed HF corpora (no network), no
third-party baselines, wall-clock per call measured around a synchronising call like the reference.

    python benchmark_metrics.py [--samples 512] [--duration 16] [--batch-sizes 1 2 4 8 16 32 64 128] [--out results]
"""
from __future__ import annotations

import argparse
import json
import os
import time

import numpy as np
import torch

from fast_speech_enhancement_metrics_b200 import PESQ, STOI
from fast_speech_enhancement_metrics_b200.synth import synth_batch

DROP_FIRST = 0.15          # the reference discards the first 15 % of the batches as warm-up


def time_metric(metric, clean, noisy, batch_size):
    times, values = [], []
    for lo in range(0, clean.shape[0] - batch_size + 1, batch_size):
        c = clean[lo:lo + batch_size].to(metric.device)
        d = noisy[lo:lo + batch_size].to(metric.device)
        t0 = time.time()
        out = metric(c, d)              # returns Python floats: synchronises like the reference's .item()
        times.append(time.time() - t0)
        values.extend(out)
    keep_from = int(len(times) * DROP_FIRST + 1)
    return {"batch_times": times[keep_from:], "values": values}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=512)
    ap.add_argument("--duration", type=float, default=16.0)
    ap.add_argument("--sample-rate", type=int, default=16000)
    ap.add_argument("--batch-sizes", type=int, nargs="+", default=[1, 2, 4, 8, 16, 32, 64, 128])
    ap.add_argument("--out", default="results")
    args = ap.parse_args()
    n = int(args.duration * args.sample_rate)
    clean, noisy, snrs = synth_batch(42, args.samples, n, fs=args.sample_rate, snr_range=(-5.0, 25.0))
    clean, noisy = torch.from_numpy(clean), torch.from_numpy(noisy)
    for bs in args.batch_sizes:
        for cls in (PESQ, STOI):
            metric = cls(sample_rate=args.sample_rate, use_gpu=True)
            record = {cls.__name__ + "_GPU": time_metric(metric, clean, noisy, bs),
                      "snrs": np.asarray(snrs).tolist(), "batch_size": bs, "sample_duration": args.duration,
                      "sample_rate": args.sample_rate, "SNR_high": 25, "SNR_low": -5}
            folder = os.path.join(args.out, "batch_size_%d" % bs)
            os.makedirs(folder, exist_ok=True)
            with open(os.path.join(folder, cls.__name__ + "_results.json"), "w") as f:
                json.dump(record, f, indent=4)
            t = record[cls.__name__ + "_GPU"]["batch_times"]
            print("%s batch %d: %.1f samples/s" % (cls.__name__, bs, bs / (sum(t) / max(len(t), 1))))


if __name__ == "__main__":
    main()
