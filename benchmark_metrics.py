#!/usr/bin/env python
"""Per-batch-size timing harness writing the reference's results schema.

The reference's `benchmark_metrics.py` (its :81-108) stores, for every batch size,
`results/batch_size_<B>/<Metric>_results.json` with one entry per implementation
(`"<Class>_GPU"`: `{"batch_times": [...], "values": [...]}`) plus `snrs`, `batch_size`,
`sample_duration`, `sample_rate`, `SNR_high`, `SNR_low`; its plotting scripts
(`benchmarking/plotting/utils.py:8-35`) read exactly these keys.  This harness produces the same
files for the two metrics of this repository so those scripts work unchanged.  Like the reference driver
(its :87-108) it writes one entry per (class, device): `"<Class>_GPU"` is this library; `"<Class>_CPU"` is the
upstream implementation's CPU path (`use_gpu=False`) when `--reference-path` points at a `fast_se_metrics`
package (default: `oracle/_ref` if it has been staged), so that `load_values` (deviation plot, its
`benchmarking/plotting/utils.py:25-35`) finds a `metric_reference` / `metric_optimized` pair
(`PESQ_CPU` / `PESQ_GPU`, `STOI_CPU` / `STOI_GPU`).  Differences from the reference driver: synthetic
speech-like audio instead of the streamed HF corpora (no network), no third-party baselines (`pesq`,
`torch_pesq`, `pystoi` are not installed), no 20 s sleeps; wall-clock per call measured around a
synchronising call like the reference.

    python benchmark_metrics.py [--samples 512] [--duration 16] [--batch-sizes 1 2 4 8 16 32 64 128] [--out results]
                                [--reference-path DIR | --no-cpu] [--cpu-samples 64]
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


def load_reference(path):
    """(PESQ, STOI) classes of an upstream `fast_se_metrics` package found under `path`, or None."""
    import importlib
    import sys
    if not path or not os.path.isdir(os.path.join(path, "fast_se_metrics")):
        return None
    if path not in sys.path:
        sys.path.insert(0, path)
    try:
        return (importlib.import_module("fast_se_metrics.PESQ").PESQ, importlib.import_module("fast_se_metrics.STOI").STOI)
    except Exception as exc:      # the harness still produces the *_GPU entries
        print("reference at %s not importable: %r" % (path, exc))
        return None


def time_metric(metric, clean, noisy, batch_size, max_samples=None):
    times, values = [], []
    total = clean.shape[0] if max_samples is None else min(clean.shape[0], max(max_samples, batch_size))
    for lo in range(0, total - batch_size + 1, batch_size):
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
    ap.add_argument("--reference-path", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle", "_ref"),
                    help="directory holding an upstream fast_se_metrics package: its CPU path is timed as <Class>_CPU")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-samples", type=int, default=64, help="samples scored by the CPU reference per batch size")
    args = ap.parse_args()
    ref = None if args.no_cpu else load_reference(args.reference_path)
    n = int(args.duration * args.sample_rate)
    clean, noisy, snrs = synth_batch(42, args.samples, n, fs=args.sample_rate, snr_range=(-5.0, 25.0))
    clean, noisy = torch.from_numpy(clean), torch.from_numpy(noisy)
    for bs in args.batch_sizes:
        for k, cls in enumerate((PESQ, STOI)):
            metric = cls(sample_rate=args.sample_rate, use_gpu=True)
            record = {cls.__name__ + "_GPU": time_metric(metric, clean, noisy, bs)}
            if ref is not None:   # the upstream CPU path on the leading samples (same order, so values pair up)
                record[cls.__name__ + "_CPU"] = time_metric(ref[k](sample_rate=args.sample_rate, use_gpu=False), clean,
                                                            noisy, bs, max_samples=args.cpu_samples)
            record.update({"snrs": np.asarray(snrs).tolist(), "batch_size": bs, "sample_duration": args.duration,
                           "sample_rate": args.sample_rate, "SNR_high": 25, "SNR_low": -5})
            folder = os.path.join(args.out, "batch_size_%d" % bs)
            os.makedirs(folder, exist_ok=True)
            with open(os.path.join(folder, cls.__name__ + "_results.json"), "w") as f:
                json.dump(record, f, indent=4)
            t = record[cls.__name__ + "_GPU"]["batch_times"]
            print("%s batch %d: %.1f samples/s" % (cls.__name__, bs, bs / (sum(t) / max(len(t), 1))))


if __name__ == "__main__":
    main()
