"""benchmark_metrics.py writes the reference's results schema: run it on a tiny sample and
read the files back with the logic of the reference's plotting loaders (benchmarking/plotting/utils.py:8-35,
restated here because /root/reference does not exist on the GPU box)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu

# the keys the reference's plots look for (benchmarking/plotting/settings.py: SAMPLES_PER_SECOND metrics and
# DEVIATION_SETTINGS with the third-party baselines replaced by the upstream CPU path)
PLOT_KEYS = ["PESQ_CPU", "PESQ_GPU", "STOI_CPU", "STOI_GPU"]
DEVIATION = [{"name": "PESQ", "metric": "PESQ", "metric_optimized": "PESQ_GPU", "metric_reference": "PESQ_CPU", "bar": 1e-3},
             {"name": "STOI", "metric": "STOI", "metric_optimized": "STOI_GPU", "metric_reference": "STOI_CPU", "bar": 1e-4},
             {"name": "STOI", "metric": "ESTOI", "metric_optimized": "STOI_GPU", "metric_reference": "STOI_CPU", "bar": 1e-4}]


def load_results(results_dir):
    """utils.py:8-22: rows of (samples per second, metric key, batch size) from every *.json below results_dir."""
    rows = []
    for dirpath, _, files in os.walk(results_dir):
        for name in files:
            if not name.endswith(".json"):
                continue
            with open(os.path.join(dirpath, name)) as f:
                data = json.load(f)
            for key, entry in data.items():
                if key in PLOT_KEYS:
                    for sps in data["batch_size"] / np.array(entry["batch_times"]):
                        rows.append((float(sps), key, data["batch_size"]))
    return rows


def load_values(results_dir, settings):
    """utils.py:25-35: the paired per-sample values of the optimised and the reference implementation."""
    with open(os.path.join(results_dir, "%s_results.json" % settings["name"])) as f:
        results = json.load(f)
    ref = [v[settings["metric"]] for v in results[settings["metric_reference"]]["values"]]
    opt = [v[settings["metric"]] for v in results[settings["metric_optimized"]]["values"]]
    return opt, ref


def test_harness_writes_the_reference_schema(tmp_path):
    out = str(tmp_path / "results")
    have_ref = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "fast_se_metrics"))
    cmd = [sys.executable, os.path.join(ROOT, "benchmark_metrics.py"), "--samples", "16", "--duration", "2",
           "--batch-sizes", "2", "8", "--out", out, "--cpu-samples", "16"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    for bs in (2, 8):
        for name in ("PESQ", "STOI"):
            with open(os.path.join(out, "batch_size_%d" % bs, "%s_results.json" % name)) as f:
                rec = json.load(f)
            for key in ("snrs", "batch_size", "sample_duration", "sample_rate", "SNR_high", "SNR_low", name + "_GPU"):
                assert key in rec, key                                    # benchmark_metrics.py:97-108 of the reference
            assert rec["batch_size"] == bs and len(rec["snrs"]) == 16
            entry = rec[name + "_GPU"]
            assert len(entry["values"]) == 16 and name in entry["values"][0]
            assert len(entry["batch_times"]) == 16 // bs - int(16 // bs * 0.15 + 1)
    rows = load_results(out)
    keys = {r[1] for r in rows}
    assert {"PESQ_GPU", "STOI_GPU"} <= keys and all(r[0] > 0 for r in rows)
    if have_ref:       # the deviation plot's loader: upstream CPU path vs this library on the same samples
        assert {"PESQ_CPU", "STOI_CPU"} <= keys
        for settings in DEVIATION:
            opt, ref = load_values(os.path.join(out, "batch_size_8"), settings)
            assert len(opt) == len(ref) == 16
            assert np.max(np.abs(np.array(opt) - np.array(ref))) <= settings["bar"], settings
