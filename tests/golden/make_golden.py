#!/usr/bin/env python
"""Generate the golden fixtures by running the UNMODIFIED reference (CPU path).

Run in the build container only (needs /root/reference, torch, torchaudio):

    python tests/golden/make_golden.py

Writes tests/golden/golden_pesq.npz, golden_stoi.npz.  Inputs are NOT stored:
every case is regenerated from its seed with
fast_speech_enhancement_metrics_b200.synth (numpy PCG64) or the explicit recipe in
`cases.py`; only the reference's outputs (and a few intermediates) are stored.
Nothing at test/bench time reads /root/reference.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import torch  # noqa: E402

torch.set_num_threads(8)

from fast_se_metrics.PESQ import PESQ  # noqa: E402
from fast_se_metrics.STOI import STOI  # noqa: E402
from torchaudio.functional import lfilter  # noqa: E402

import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("golden_cases", os.path.join(HERE, "cases.py"))
_cases = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_cases)   # (the reference ships its own `tests` package, so import by path)
pesq_cases, stoi_cases, pesq_rate_cases, lsd_cases, sdr_cases = (
    _cases.pesq_cases, _cases.stoi_cases, _cases.pesq_rate_cases, _cases.lsd_cases, _cases.sdr_cases)
lsd_rate_cases, sdr_rate_cases = _cases.lsd_rate_cases, _cases.sdr_rate_cases


def run_pesq():
    out = {}
    metric = PESQ(16000, use_gpu=False)
    torch.set_float32_matmul_precision("highest")
    for name, (clean, deg, lengths) in pesq_cases().items():
        c = torch.from_numpy(clean)
        d = torch.from_numpy(deg)
        if lengths is None:
            res = [r["PESQ"] for r in metric(c, d)]
        else:  # per-item slices: legitimate by batch invariance
            res = [metric(c[i:i + 1, :n], d[i:i + 1, :n])[0]["PESQ"] for i, n in enumerate(lengths)]
        out[name] = np.asarray(res, np.float64)
        print(name, out[name])
    for name, (clean, deg, lengths, fs) in pesq_rate_cases().items():
        m = PESQ(fs, use_gpu=False)
        c = torch.from_numpy(clean)
        d = torch.from_numpy(deg)
        if lengths is None:
            res = [r["PESQ"] for r in m(c, d)]
        else:
            res = [m(c[i:i + 1, :n], d[i:i + 1, :n])[0]["PESQ"] for i, n in enumerate(lengths)]
        out["rate/" + name] = np.asarray(res, np.float64)
        print(name, out["rate/" + name])
    # stage taps for the first "speech2s" item: band power (float32 IIR as the reference
    # runs it) and the Bark-band spectrogram of the clean signal
    clean, deg, _ = pesq_cases()["speech2s"]
    c = torch.from_numpy(clean[:1])
    d = torch.from_numpy(deg[:1])
    ce, de = metric.equalize_ranges(c, d)
    both = torch.cat([ce, de], 0)
    filt = lfilter(both, metric.power_filter[1], metric.power_filter[0], clamp=False)
    out["tap_band_power_f32"] = filt.square().sum(1).numpy().astype(np.float64)
    filt64 = lfilter(both.double(), metric.power_filter[1].double(), metric.power_filter[0].double(), clamp=False)
    out["tap_band_power_f64"] = filt64.square().sum(1).numpy()
    out["tap_bark"] = metric.get_bark_bands(both.clone()).numpy()          # [2, T, 49] float64
    out["tap_hann512"] = metric.to_spec.window.numpy()
    out["tap_power_filter"] = metric.power_filter.numpy()                   # [2, 11] float32 (b, a)
    # error behaviour
    try:
        metric(torch.randn(1, 5120), torch.randn(1, 5120))    # T = 19 < 20 frames
        out["err_too_short"] = np.array(0)
    except RuntimeError:
        out["err_too_short"] = np.array(1)
    np.savez_compressed(os.path.join(HERE, "golden_pesq.npz"), **out)


def run_stoi():
    out = {}
    for name, (clean, deg, lengths, fs) in stoi_cases().items():
        metric = STOI(fs, use_gpu=False)
        c = torch.from_numpy(clean)
        d = torch.from_numpy(deg)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if lengths is None:
                res = metric(c, d)
            else:
                res = []
                for i, n in enumerate(lengths):
                    try:
                        res.append(metric(c[i:i + 1, :n], d[i:i + 1, :n])[0])
                    except TypeError:      # whole batch without segments (STOI.py:163-165,205)
                        res.append({"STOI": float("nan"), "ESTOI": float("nan")})
        out[name + "/stoi"] = np.asarray([r["STOI"] for r in res], np.float64)
        out[name + "/estoi"] = np.asarray([r["ESTOI"] for r in res], np.float64)
        # silent-frame mask and K of every item (the bit-exact decisions), on the 10 kHz signal
        ks, masks = [], []
        for i in range(c.shape[0]):
            n = c.shape[1] if lengths is None else lengths[i]
            ci = metric.prepare_audio(c[i, :n])
            if ci.shape[1] < 256:
                ks.append(0)
                masks.append(np.zeros(0, np.uint8))
                continue
            fr = ci.unfold(1, 256, 128) * metric.window
            e = 20 * torch.log10(torch.norm(fr, dim=2) + 1e-9)
            m = (torch.amax(e, dim=1, keepdim=True) - 40 - e) < 0
            ks.append(int(m.sum()))
            masks.append(np.packbits(m[0].numpy().astype(np.uint8)))
        out[name + "/K"] = np.asarray(ks, np.int64)
        out[name + "/mask_bits"] = np.concatenate(masks)
        out[name + "/mask_nbytes"] = np.asarray([len(m) for m in masks], np.int64)
        print(name, out[name + "/stoi"], out[name + "/estoi"], out[name + "/K"])
    # stage taps
    metric = STOI(16000, use_gpu=False)
    clean, deg, _, _ = stoi_cases()["speech16k_3s"]
    c10 = metric.prepare_audio(torch.from_numpy(clean[:1]))
    d10 = metric.prepare_audio(torch.from_numpy(deg[:1]))
    out["tap_resampled_clean"] = c10[0, :4000].numpy()
    out["tap_resample_kernel"] = metric.resampler.kernel.numpy()[:, 0, :]
    out["tap_window"] = metric.window.numpy()
    out["tap_obm"] = metric.octave_band_matrix.numpy()
    cs, ds, ls = metric.remove_silent_frames(c10, d10)
    spec = metric.stft(torch.cat([cs, ds], 0), torch.cat([ls, ls], 0))
    tob = torch.sqrt(torch.matmul(metric.octave_band_matrix, spec))
    u = int(1 + (ls[0].item() - 512) // 128)
    out["tap_tob"] = tob[:, :, :u].numpy()            # [2, 15, U]
    # error behaviour: whole batch without segments -> warning + TypeError
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            STOI(10000, False)(torch.randn(2, 3000), torch.randn(2, 3000))
        out["err_no_segments"] = np.array(0)
    except TypeError:
        out["err_no_segments"] = np.array(1)
    np.savez_compressed(os.path.join(HERE, "golden_stoi.npz"), **out)


def run_lsd():
    from fast_se_metrics.LSD import LSD
    out = {}
    metric = LSD(16000, use_gpu=False)
    for name, (clean, deg, lengths) in lsd_cases().items():
        c, d = torch.from_numpy(clean), torch.from_numpy(deg)
        if lengths is None:
            res = [r["LSD"] for r in metric(c, d)]
        else:
            res = [metric(c[i:i + 1, :n], d[i:i + 1, :n])[0]["LSD"] for i, n in enumerate(lengths)]
        out[name] = np.asarray(res, np.float64)
        print("LSD", name, out[name])
    for name, (clean, deg, lengths, fs) in lsd_rate_cases().items():      # resample-on-ingest (base.py:19-20)
        m = LSD(fs, use_gpu=False)
        c, d = torch.from_numpy(clean), torch.from_numpy(deg)
        if lengths is None:
            res = [r["LSD"] for r in m(c, d)]
        else:
            res = [m(c[i:i + 1, :n], d[i:i + 1, :n])[0]["LSD"] for i, n in enumerate(lengths)]
        out["rate/" + name] = np.asarray(res, np.float64)
        print("LSD rate", name, out["rate/" + name])
    np.savez_compressed(os.path.join(HERE, "golden_lsd.npz"), **out)


def run_sdr():
    from fast_se_metrics.SDR import SDR
    out = {}
    metric = SDR(16000, use_gpu=False)
    for name, (clean, deg, lengths) in sdr_cases().items():
        c, d = torch.from_numpy(clean), torch.from_numpy(deg)
        if lengths is None:
            res = [r["SDR"] for r in metric(c, d)]
        else:
            res = [metric(c[i:i + 1, :n], d[i:i + 1, :n])[0]["SDR"] for i, n in enumerate(lengths)]
        out[name] = np.asarray(res, np.float64)
        print("SDR", name, out[name])
    for name, (clean, deg, lengths, fs) in sdr_rate_cases().items():      # resample-on-ingest (base.py:19-20)
        m = SDR(fs, use_gpu=False)
        c, d = torch.from_numpy(clean), torch.from_numpy(deg)
        if lengths is None:
            res = [r["SDR"] for r in m(c, d)]
        else:
            res = [m(c[i:i + 1, :n], d[i:i + 1, :n])[0]["SDR"] for i, n in enumerate(lengths)]
        out["rate/" + name] = np.asarray(res, np.float64)
        print("SDR rate", name, out["rate/" + name])
    np.savez_compressed(os.path.join(HERE, "golden_sdr.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["pesq", "stoi", "lsd", "sdr"]      # e.g. `make_golden.py lsd sdr` regenerates two files
    for name in which:
        {"pesq": run_pesq, "stoi": run_stoi, "lsd": run_lsd, "sdr": run_sdr}[name]()
    print("golden fixtures written to", HERE)
