"""Host-side constant design (the one-off work of the reference's constructors).

PESQ.__init__ (PESQ.py:55-90), BarkFilterBank.__init__ (utils/bark.py:129-164),
Loudness.__init__ (utils/loudness.py:42-46), STOI.__init__ (STOI.py:11-47) and the
torchaudio Resample kernel built by BaseMetric.__init__ (base.py:13).  The results
are packed into the plain-C design structs of include/fsem.h.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch
from scipy.signal import butter, lfilter, residuez

from . import p862_tables as P862

NB = 49
SECTIONS = 5


class PesqDesign(C.Structure):
    """Mirror of fsem_pesq_design_t (include/fsem.h)."""
    _fields_ = [
        ("bp_direct", C.c_float),
        ("bp_c0", C.c_float * SECTIONS), ("bp_c1", C.c_float * SECTIONS),
        ("bp_a1", C.c_float * SECTIONS), ("bp_a2", C.c_float * SECTIONS),
        ("pre_b", C.c_float * 3), ("pre_a", C.c_float * 2),
        ("taper", C.c_float * 15),
        ("warmup", C.c_int32),
        ("hann", C.c_float * 512),
        ("band_first_bin", C.c_int32 * NB), ("band_num_bins", C.c_int32 * NB),
        ("pow_dens", C.c_float * NB), ("thresh", C.c_float * NB),
        ("zwicker_exp", C.c_float * NB), ("width_bark", C.c_float * NB),
        ("sl", C.c_float),
        ("rs_orig", C.c_int32), ("rs_neu", C.c_int32), ("rs_width", C.c_int32), ("rs_ntaps", C.c_int32),
        ("rs_taps", C.POINTER(C.c_float)),
    ]


class StoiDesign(C.Structure):
    """Mirror of fsem_stoi_design_t (include/fsem.h)."""
    _fields_ = [
        ("orig", C.c_int32), ("neu", C.c_int32), ("width", C.c_int32), ("ntaps", C.c_int32),
        ("taps", C.POINTER(C.c_float)),
        ("window", C.c_float * 256),
        ("band_lo", C.c_int32 * 15), ("band_hi", C.c_int32 * 15),
        ("clip", C.c_float), ("dyn_range", C.c_float),
    ]


def _fill(dst, values):
    for i, v in enumerate(values):
        dst[i] = v


# ------------------------------------------------------------------------------------------- PESQ
def power_filter_f32():
    """The reference's level-alignment filter exactly as it stores it: scipy
    butter(5, [325, 3250], fs=16000, 'band') cast to float32 (PESQ.py:80-81)."""
    b, a = butter(5, [325, 3250], fs=16000, btype="band")
    return np.asarray(b, np.float32), np.asarray(a, np.float32)


def parallel_sections(b32: np.ndarray, a32: np.ndarray):
    """Partial-fraction (parallel) form of the float32-rounded transfer function:
        H(z) = k + sum_s (c0_s + c1_s z^-1) / (1 + a1_s z^-1 + a2_s z^-2).
    The direct form the reference runs in float32 is ill-conditioned (order 10); the
    parallel form of the SAME rounded polynomials is its exact value in float32-safe
    arithmetic, 22 flops/sample, and its five sections are independent (ILP).
    Returns (k, sections[5, 4] = c0, c1, a1, a2) in float64."""
    r, p, k = residuez(b32.astype(np.float64), a32.astype(np.float64))
    if len(k) != 1:
        raise ValueError("expected a proper-plus-constant transfer function")
    cplx = [i for i in range(len(p)) if p[i].imag > 1e-9]
    real = [i for i in range(len(p)) if abs(p[i].imag) <= 1e-9]
    if 2 * len(cplx) + len(real) != len(p) or len(real) % 2:
        raise ValueError("unexpected pole structure")
    secs = []
    for i in cplx:      # r/(1-p z^-1) + r*/(1-p* z^-1)
        secs.append((2 * r[i].real, -2 * (r[i] * np.conj(p[i])).real, -2 * p[i].real, abs(p[i]) ** 2))
    real.sort(key=lambda i: p[i].real)
    for j in range(0, len(real), 2):   # two real poles share one second-order section
        i1, i2 = real[j], real[j + 1]
        secs.append(((r[i1] + r[i2]).real, -(r[i1] * p[i2] + r[i2] * p[i1]).real,
                     -(p[i1] + p[i2]).real, (p[i1] * p[i2]).real))
    if len(secs) != SECTIONS:
        raise ValueError("expected %d second-order sections, got %d" % (SECTIONS, len(secs)))
    secs = np.asarray(secs, np.float64)
    # self-check against the direct form in float64
    imp = np.zeros(2048)
    imp[0] = 1.0
    h_ref = lfilter(b32.astype(np.float64), a32.astype(np.float64), imp)
    h_par = k[0].real * imp
    for c0, c1, a1, a2 in secs:
        h_par = h_par + lfilter([c0, c1], [1.0, a1, a2], imp)
    if np.max(np.abs(h_par - h_ref)) > 1e-9 * np.max(np.abs(h_ref)):
        raise ValueError("parallel-form decomposition does not reproduce the filter")
    return float(k[0].real), secs


def _decay_samples(h: np.ndarray, rel: float) -> int:
    tail = np.maximum.accumulate(np.abs(h[::-1]))[::-1]
    idx = np.nonzero(tail <= rel * np.abs(h).max())[0]
    return int(idx[0]) if len(idx) else len(h)


def pesq_design(sample_rate: int = 16000, target_rate: int = 16000):
    """Returns (PesqDesign, resampling taps kept alive by the caller or None)."""
    d = PesqDesign()
    taps = None
    if sample_rate != target_rate:
        taps, width, o, nw = sinc_hann_kernel(sample_rate, target_rate)
        d.rs_orig, d.rs_neu, d.rs_width, d.rs_ntaps = o, nw, width, taps.shape[1]
        d.rs_taps = taps.ctypes.data_as(C.POINTER(C.c_float))
    else:
        d.rs_orig, d.rs_neu, d.rs_width, d.rs_ntaps = 1, 1, 0, 0
        d.rs_taps = None
    b32, a32 = power_filter_f32()
    k, secs = parallel_sections(b32, a32)
    d.bp_direct = k
    _fill(d.bp_c0, secs[:, 0]); _fill(d.bp_c1, secs[:, 1])
    _fill(d.bp_a1, secs[:, 2]); _fill(d.bp_a2, secs[:, 3])
    pre_b = np.array([2.740826, -5.4816519, 2.740826], np.float32)      # PESQ.py:85
    pre_a = np.array([1.0, -1.9444777, 0.94597794], np.float32)        # PESQ.py:86
    _fill(d.pre_b, pre_b); _fill(d.pre_a, pre_a[1:])
    _fill(d.taper, (torch.linspace(0, 15, 16)[1:] / 16.0).tolist())     # PESQ.py:90
    imp = np.zeros(8192)
    imp[0] = 1.0
    w = max(_decay_samples(lfilter(b32.astype(np.float64), a32.astype(np.float64), imp), 2e-9),
            _decay_samples(lfilter(pre_b.astype(np.float64), pre_a.astype(np.float64), imp), 2e-9))
    d.warmup = int(min(max(64 * math.ceil(w / 64), 256), 4096))
    _fill(d.hann, torch.hann_window(512, dtype=torch.float32).tolist())  # Spectrogram window (PESQ.py:63-71)
    first = np.concatenate([[0], np.cumsum(P862.BINS_PER_BAND)[:-1]])    # bark.py:137-148
    _fill(d.band_first_bin, first.tolist()); _fill(d.band_num_bins, P862.BINS_PER_BAND.tolist())
    _fill(d.pow_dens, (P862.POW_DENS_CORRECTION * P862.SP).tolist())     # bark.py:132
    _fill(d.thresh, P862.ABS_THRESH_POWER.tolist())                      # loudness.py:43
    exps = np.clip(6.0 / (P862.CENTRE_BARK + 2.0), 1.0, 2.0) ** 0.15 * P862.ZWICKER_POWER   # loudness.py:45-46
    _fill(d.zwicker_exp, exps.tolist())
    _fill(d.width_bark, P862.WIDTH_BARK.tolist())
    d.sl = P862.SL
    return d, taps


# ------------------------------------------------------------------------------------------- STOI
def sinc_hann_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """torchaudio's sinc_interp_hann resampling kernel (what Resample(orig, new) builds,
    base.py:13): float64 maths (with torchaudio's float32 phase offsets), float32 taps.
    Returns (taps[new, 2*width+orig], width, orig_reduced, new_reduced)."""
    g = math.gcd(int(orig_freq), int(new_freq))
    o, nw = int(orig_freq) // g, int(new_freq) // g
    base = min(o, nw) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    idx = np.arange(-width, width + o, dtype=np.float64)[None, :] / o
    phase = (np.arange(0, -nw, -1).astype(np.float32) / np.float32(nw)).astype(np.float64)
    t = (phase[:, None] + idx) * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2.0) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        taps = np.where(t == 0, 1.0, np.sin(t) / t) * window * (base / o)
    return np.ascontiguousarray(taps, np.float32), width, o, nw


def third_octave_bins(fs: int = 10000, n_fft: int = 512, num_bands: int = 15, f_min: float = 150.0):
    """[lo, hi) FFT-bin range of every one-third-octave band (STOI.py:26-47)."""
    freqs = np.linspace(0, fs // 2, n_fft // 2 + 1)
    k = np.arange(num_bands, dtype=np.float64)
    lo = [int(np.argmin(np.abs(freqs - f))) for f in f_min * 2.0 ** ((2 * k - 1) / 6)]
    hi = [int(np.argmin(np.abs(freqs - f))) for f in f_min * 2.0 ** ((2 * k + 1) / 6)]
    return lo, hi


def stoi_design(sample_rate: int, target_rate: int = 10000):
    """Returns (StoiDesign, taps array kept alive by the caller)."""
    d = StoiDesign()
    taps = None
    if sample_rate != target_rate:
        taps, width, o, nw = sinc_hann_kernel(sample_rate, target_rate)
        d.orig, d.neu, d.width, d.ntaps = o, nw, width, taps.shape[1]
        d.taps = taps.ctypes.data_as(C.POINTER(C.c_float))
    else:
        d.orig, d.neu, d.width, d.ntaps = 1, 1, 0, 0
        d.taps = None
    _fill(d.window, torch.hann_window(257, dtype=torch.float32)[1:].tolist())   # STOI.py:24
    lo, hi = third_octave_bins(target_rate)
    _fill(d.band_lo, lo); _fill(d.band_hi, hi)
    d.clip = 1.0 + 10.0 ** (15.0 / 20.0)      # beta = -15 dB (STOI.py:21, 137-138)
    d.dyn_range = 40.0                        # STOI.py:22
    return d, taps
