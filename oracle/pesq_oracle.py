"""Float64 numpy restatement of the reference's wide-band PESQ path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never imported by the product.

Reference: /root/reference/fast_se_metrics/PESQ.py, utils/bark.py, utils/loudness.py
(cited per function as file:line).  Third-party arithmetic underneath the
reference that is restated here from its published definition:
  * torchaudio.functional.lfilter (torchaudio 2.8.0 in poetry.lock, 2.11.0 in
    the image): direct-form IIR  y[t] = sum b[k] x[t-k] - sum a[k] y[t-k], zero
    initial state, computed by the reference in float32.  Here: the SAME
    float32-rounded coefficients, evaluated in float64 (scipy.signal.lfilter).
    The reference's own float32 recursion is ill-conditioned (order 10) and adds
    ~1e-4 of noise to the PESQ score; this oracle is the noise-free value.
  * torchaudio.transforms.Spectrogram(n_fft=512, hop=256, hann, power=2,
    center=False): |rfft(hann_periodic(512) * frame)|^2.
  * scipy.signal.butter (design only).

Everything is per item (the reference is batch-invariant: no cross-item op).
"""
from __future__ import annotations

import numpy as np
from scipy.signal import butter, lfilter

from . import tables as T

N_FFT = 512
HOP = 256
N_BANDS = 49


def power_filter_coeffs():
    """PESQ.py:80-81 -- Butterworth band-pass 325..3250 Hz, order 5 (10th-order IIR),
    coefficients rounded to float32 exactly as the reference stores them."""
    b, a = butter(5, [325, 3250], fs=16000, btype="band")
    return (np.asarray(b, np.float32).astype(np.float64),
            np.asarray(a, np.float32).astype(np.float64))


def pre_filter_coeffs():
    """PESQ.py:84-88 -- pre-emphasis biquad constants (float32 in the reference)."""
    b = np.array([2.740826, -5.4816519, 2.740826], np.float32).astype(np.float64)
    a = np.array([1.0, -1.9444777, 0.94597794], np.float32).astype(np.float64)
    return b, a


_PB, _PA = power_filter_coeffs()
_EB, _EA = pre_filter_coeffs()
_TAPER = (np.arange(1, 16, dtype=np.float32) / np.float32(16.0)).astype(np.float64)  # PESQ.py:90
_HANN = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(N_FFT) / N_FFT)                   # periodic Hann (PESQ.py:63-71)
_BAND_EDGES = np.concatenate([[0], np.cumsum(T.BINS_PER_BAND)])                       # bark.py:137-148
_POW_DENS = T.POWER_DENSITY_CORRECTION * T.SP_16K                                     # bark.py:132
_THR = T.ABS_THRESHOLD_POWER                                                          # loudness.py:43
_EXP = np.clip(6.0 / (T.BAND_CENTRE_BARK + 2.0), 1.0, 2.0) ** 0.15 * T.ZWICKER_POWER  # loudness.py:45-46
_W = T.BAND_WIDTH_BARK
_WTOT = _W[1:].sum()                                                                  # bark.py:164


def band_power(x: np.ndarray) -> float:
    """PESQ.py:94-97 -- sum of squares of the band-passed signal (before the
    /(n+5120)/1.04684 normalisation)."""
    y = lfilter(_PB, _PA, x)
    return float(np.dot(y, y))


def align_level(x: np.ndarray) -> np.ndarray:
    """PESQ.py:92-102."""
    n = x.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        power = np.float64(band_power(x)) / (n + 5120) / 1.04684     # numpy scalar: 1e7/0 -> inf, 0*inf -> nan
        return x * np.sqrt(np.float64(1e7) / power)


def pre_emphasize(x: np.ndarray) -> np.ndarray:
    """PESQ.py:104-113 -- 15-sample linear tapers at both ends, then the biquad."""
    x = x.copy()
    x[:15] *= _TAPER
    x[-15:] *= _TAPER[::-1]
    return lfilter(_EB, _EA, x)


def num_frames(n: int) -> int:
    """PESQ.py:128-133 -- right-pad by n % 256 zeros, frames of 512 hop 256, no centring."""
    return 1 + (n + n % HOP - N_FFT) // HOP


def bark_bands(x: np.ndarray) -> np.ndarray:
    """PESQ.py:123-140 + bark.py:203-204 -- [T, 49] Bark-band power densities."""
    z = pre_emphasize(align_level(x))
    n = z.shape[0]
    z = np.concatenate([z, np.zeros(n % HOP)])
    t = 1 + (z.shape[0] - N_FFT) // HOP
    if t < 1:
        raise RuntimeError("signal shorter than one frame")
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(t)[:, None]
    spec = np.abs(np.fft.rfft(z[idx] * _HANN[None, :], axis=1)) ** 2     # [T, 257]
    spec[:, 0] = 0.0                                                      # PESQ.py:136
    bands = np.add.reduceat(spec[:, :256], _BAND_EDGES[:-1], axis=1)      # bins 0..255 only (bark.py:203)
    return bands * _POW_DENS[None, :]


def audible_frame_power(bands: np.ndarray, factor: float) -> np.ndarray:
    """loudness.py:48-53."""
    return np.sum(bands * (bands > _THR[None, :] * factor), axis=1)


def mean_audible_band_power(bands: np.ndarray, silent: np.ndarray) -> np.ndarray:
    """loudness.py:55-60 -- mean over ALL frames."""
    mask = (bands > _THR[None, :] * 100.0) & (~silent)[:, None]
    return np.mean(bands * mask, axis=0)


def loudness(p: np.ndarray) -> np.ndarray:
    """loudness.py:62-67 -- Zwicker's law."""
    with np.errstate(invalid="ignore"):
        l = (2.0 * _THR) ** _EXP * ((0.5 + 0.5 * p / _THR) ** _EXP - 1.0)
    l = np.where(p <= _THR, 0.0, l)
    return l * T.SL_16K


def weighted_norm(x: np.ndarray, p: float) -> np.ndarray:
    """bark.py:169-184 -- band 0 is excluded."""
    v = np.abs(_W[None, 1:] * x[:, 1:] / _WTOT ** (1.0 / p))
    return _WTOT * np.sum(v ** p, axis=1) ** (1.0 / p)


def overlapping_sums(d: np.ndarray) -> float:
    """PESQ.py:168-172 -- windows of 20 frames hop 10: L6 inside, L2 across."""
    t = d.shape[0]
    if t < 20:
        raise RuntimeError("fewer than 20 frames")
    w = (t - 20) // 10 + 1
    idx = np.arange(20)[None, :] + 10 * np.arange(w)[:, None]
    psqm = np.mean(d[idx] ** 6, axis=1) ** (1.0 / 6.0)
    return float(np.sqrt(np.mean(psqm ** 2)))


def pesq_item(clean: np.ndarray, deg: np.ndarray, taps: dict | None = None) -> float:
    """PESQ.py:174-245 for one (clean, degraded) pair of float32 signals.

    `taps`, when given, receives intermediate results (stage taps for kernel tests).
    """
    c = np.asarray(clean, np.float64)
    d = np.asarray(deg, np.float64)
    assert c.ndim == 1 and c.shape == d.shape
    with np.errstate(divide="ignore", invalid="ignore"):
        m = max(np.abs(c).max(), np.abs(d).max())            # PESQ.py:115-121
        c = c / m
        d = d / m
        bc = bark_bands(c)
        bd = bark_bands(d)

        # PESQ.py:142-166
        silent = audible_frame_power(bc, 1e2) < 1e7
        ratio = (mean_audible_band_power(bd, silent) + 1000.0) / (mean_audible_band_power(bc, silent) + 1000.0)
        ratio = np.clip(ratio, 0.01, 100.0)
        bc_eq = ratio[None, :] * bc
        fr = (audible_frame_power(bc_eq, 1.0) + 5e3) / (audible_frame_power(bd, 1.0) + 5e3)
        fr_s = fr.copy()
        fr_s[1:] = 0.8 * fr[1:] + 0.2 * fr[:-1]               # non-recursive (RHS evaluated first)
        fr_s = np.clip(fr_s, 3e-4, 5.0)
        bd_eq = fr_s[:, None] * bd

        lc = loudness(bc_eq)
        ld = loudness(bd_eq)

        # PESQ.py:197-200
        dead = 0.25 * np.minimum(lc, ld)
        dist = ld - lc
        dist = np.sign(dist) * np.maximum(np.abs(dist) - dead, 0.0)

        sym = np.maximum(weighted_norm(dist, 2.0), 1e-20)     # PESQ.py:210-211
        scale = ((bd_eq + 50.0) / (bc_eq + 50.0)) ** 1.2      # PESQ.py:214-216
        scale = np.where(scale < 3.0, 0.0, scale)
        scale = np.minimum(scale, 12.0)
        asym = np.maximum(weighted_norm(dist * scale, 1.0), 1e-20)

        weight = ((audible_frame_power(bc_eq, 1.0) + 1e5) / 1e7) ** 0.04   # PESQ.py:222-224
        sym = np.minimum(sym / weight, 45.0)
        asym = np.minimum(asym / weight, 45.0)
        # torch.clamp propagates NaN; np.minimum/maximum do too.

        d_sym = overlapping_sums(sym)
        d_asym = overlapping_sums(asym)
        mos = 4.5 - 0.1 * d_sym - 0.0309 * d_asym             # PESQ.py:240
        mos = 0.999 + 4.0 / (1.0 + np.exp(-1.3669 * mos + 3.8224))  # PESQ.py:243
    if taps is not None:
        taps.update(bark_clean=bc, bark_deg=bd, ratio=ratio, frame_ratio=fr_s, sym=sym, asym=asym,
                    d_sym=d_sym, d_asym=d_asym)
    return float(mos)


def pesq_batch(clean: np.ndarray, deg: np.ndarray, lengths=None, sample_rate: int = 16000) -> np.ndarray:
    """Per-item loop over a [B, n] batch; `lengths` slices each row (the reference
    called on x[i, :len_i]; legitimate by batch invariance).  sample_rate != 16000:
    resample-on-ingest first (base.py:19-20, torchaudio sinc-Hann kernel)."""
    from .stoi_oracle import resample
    clean = np.atleast_2d(clean)
    deg = np.atleast_2d(deg)
    out = np.empty(clean.shape[0], np.float64)
    for i in range(clean.shape[0]):
        n = clean.shape[1] if lengths is None else int(lengths[i])
        out[i] = pesq_item(resample(clean[i, :n], sample_rate, 16000), resample(deg[i, :n], sample_rate, 16000))
    return out
