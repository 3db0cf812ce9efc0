"""Float64 numpy restatement of the reference's STOI / ESTOI path (plus the
16 kHz -> 10 kHz resample-on-ingest of base.py).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never imported by the product.

Reference: /root/reference/fast_se_metrics/STOI.py, base.py (file:line per function).
Third-party arithmetic restated from its published definition:
  * torchaudio.transforms.Resample (sinc_interp_hann, lowpass_filter_width=6,
    rolloff=0.99; torchaudio 2.8.0 pinned / 2.11.0 in the image): polyphase FIR
    whose taps are computed in float64 and stored as float32; the reference applies
    them with a float32 conv1d.  Here: same float32 taps, float64 accumulation.
  * torch.stft(n_fft=512, win_length=256, hop=128, center=False): the 256-sample
    window is zero-padded to 512 centred (128 zeros either side).

The route is deliberately LITERAL (frame -> mask -> overlap-add -> STFT -> stack of
30-frame segments), so that the CUDA path's algebraic shortcuts (no overlap-add
buffer, sliding segments) are checked against an independent formulation.
The 1e-12 * randn term of STOI.normalize (STOI.py:116) is omitted: it is below
float32 resolution for any non-constant row and makes the reference itself
non-deterministic for constant rows.
"""
from __future__ import annotations

import math

import numpy as np

FS = 10000
WIN = 256
HOP = 128
N_FFT = 512
N_BANDS = 15
SEG = 30
DYN_RANGE = 40.0
CLIP = 1.0 + 10.0 ** (15.0 / 20.0)     # STOI.py:137  (beta = -15 dB)


def resample_kernel(orig: int, new: int, lowpass_filter_width: int = 6, rolloff: float = 0.99):
    """torchaudio _get_sinc_resample_kernel (sinc_interp_hann), float64 maths, float32 taps.
    Returns (taps[new, 2*width+orig] float32, width, orig_reduced, new_reduced).  base.py:13."""
    g = math.gcd(int(orig), int(new))
    o, nw = int(orig) // g, int(new) // g
    base = min(o, nw) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    idx = np.arange(-width, width + o, dtype=np.float64)[None, :] / o
    # torchaudio divides an int64 arange by new_freq -> float32 phase offsets, then adds the float64 idx
    phase = (np.arange(0, -nw, -1).astype(np.float32) / np.float32(nw)).astype(np.float64)
    t = phase[:, None] + idx
    t = t * base
    t = np.clip(t, -lowpass_filter_width, lowpass_filter_width)
    window = np.cos(t * math.pi / lowpass_filter_width / 2.0) ** 2
    t = t * math.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    k = k * window * (base / o)
    return k.astype(np.float32), width, o, nw


def resample(x: np.ndarray, orig: int, new: int) -> np.ndarray:
    """torchaudio _apply_sinc_resample_kernel: y[new*k+p] = sum_j taps[p,j] * xpad[orig*k+j],
    xpad = width zeros | x | width+orig zeros; output cut to ceil(new*n/orig).  base.py:19-20.
    Result rounded to float32 (the reference's conv1d output dtype)."""
    if orig == new:
        return np.asarray(x, np.float32)
    taps, width, o, nw = resample_kernel(orig, new)
    x = np.asarray(x, np.float64)
    n = x.shape[0]
    xpad = np.concatenate([np.zeros(width), x, np.zeros(width + o)])
    ntap = taps.shape[1]
    nblk = (xpad.shape[0] - ntap) // o + 1
    idx = o * np.arange(nblk)[:, None] + np.arange(ntap)[None, :]
    y = xpad[idx] @ taps.astype(np.float64).T            # [nblk, new]
    y = y.reshape(-1)[: int(math.ceil(nw * n / o))]
    return y.astype(np.float32)


def octave_band_edges():
    """STOI.py:26-47 -- 15 one-third-octave bands as [lo, hi) FFT-bin ranges (512-pt @ 10 kHz)."""
    freqs = np.linspace(0, FS // 2, N_FFT // 2 + 1)
    k = np.arange(N_BANDS, dtype=np.float64)
    f_lo = 150.0 * 2.0 ** ((2 * k - 1) / 6)
    f_hi = 150.0 * 2.0 ** ((2 * k + 1) / 6)
    lo = np.array([np.argmin(np.abs(freqs - f)) for f in f_lo])
    hi = np.array([np.argmin(np.abs(freqs - f)) for f in f_hi])
    return lo, hi


_WINDOW32 = None


def window32() -> np.ndarray:
    """STOI.py:24 -- hann_window(257)[1:] in float32 (torch computes the periodic Hann
    in float32; regenerated here in float64 and rounded -- checked against the
    golden fixture)."""
    global _WINDOW32
    if _WINDOW32 is None:
        k = np.arange(1, WIN + 1, dtype=np.float64)
        _WINDOW32 = (0.5 - 0.5 * np.cos(2.0 * np.pi * k / (WIN + 1))).astype(np.float32)
    return _WINDOW32


def silent_frame_mask(clean10k: np.ndarray, window: np.ndarray | None = None):
    """STOI.py:92-102 -- the bit-exact decision.  The frame norm is evaluated in
    float64 from the float32 products (w*x rounded to float32 as the reference
    does), rounded to float32, and the remaining ops (+1e-9, log10, *20, max,
    -40, -E, <0) are done in float32 in the reference's order.  Returns (mask, E)."""
    w = window32() if window is None else window
    x = np.asarray(clean10k, np.float32)
    t0 = (x.shape[0] - WIN) // HOP + 1
    if t0 < 1:
        return np.zeros(0, bool), np.zeros(0, np.float32)
    idx = HOP * np.arange(t0)[:, None] + np.arange(WIN)[None, :]
    fr = (x[idx] * w[None, :]).astype(np.float32)                 # float32 products
    nrm = np.sqrt(np.sum(fr.astype(np.float64) ** 2, axis=1)).astype(np.float32)
    e = (np.float32(20.0) * np.log10((nrm + np.float32(1e-9)).astype(np.float64)).astype(np.float32)).astype(np.float32)
    mx = e.max()
    mask = ((mx - np.float32(DYN_RANGE)).astype(np.float32) - e).astype(np.float32) < 0
    return mask, e


def remove_silent_frames(clean: np.ndarray, deg: np.ndarray):
    """STOI.py:88-111 + overlap_and_add STOI.py:71-86: returns (clean_ola, deg_ola, K)."""
    w = window32().astype(np.float64)
    mask, _ = silent_frame_mask(clean)
    k = int(mask.sum())
    out = []
    for x in (clean, deg):
        x = np.asarray(x, np.float64)
        t0 = mask.shape[0]
        idx = HOP * np.arange(t0)[:, None] + np.arange(WIN)[None, :]
        fr = (x[idx] * w[None, :])[mask] if t0 else np.zeros((0, WIN))
        y = np.zeros((k + 1) * HOP)
        for j in range(k):
            y[j * HOP: j * HOP + WIN] += fr[j]
        out.append(y)
    return out[0], out[1], k


def third_octave(y: np.ndarray) -> np.ndarray:
    """STOI.py:49-69 + 121-125 -- power STFT (window zero-padded to 512, centred) and
    sqrt of the third-octave band sums: returns tob[15, U]."""
    w = window32().astype(np.float64)
    wpad = np.concatenate([np.zeros(128), w, np.zeros(128)])
    u = 1 + (y.shape[0] - N_FFT) // HOP
    if u < 1:
        return np.zeros((N_BANDS, 0))
    idx = HOP * np.arange(u)[:, None] + np.arange(N_FFT)[None, :]
    spec = np.abs(np.fft.rfft(y[idx] * wpad[None, :], axis=1)) ** 2      # [U, 257]
    lo, hi = octave_band_edges()
    tob = np.stack([spec[:, a:b].sum(axis=1) for a, b in zip(lo, hi)], axis=0)
    return np.sqrt(tob)


def _normalize(x: np.ndarray, axis: int) -> np.ndarray:
    """STOI.py:113-119 without the 1e-12 noise."""
    x = x - x.mean(axis=axis, keepdims=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        return x / np.linalg.norm(x, axis=axis, keepdims=True)


def stoi_item_10k(clean: np.ndarray, deg: np.ndarray, taps: dict | None = None):
    """STOI.py:153-198 for one pair already at 10 kHz.  Returns (stoi, estoi, K);
    K <= 31 gives (nan, nan, K) -- the reference's 0/0 for an item without segments."""
    yc, yd, k = remove_silent_frames(clean, deg)
    xc = third_octave(yc)
    xd = third_octave(yd)
    m = max(k - 31, 0)
    if taps is not None:
        taps.update(tob_clean=xc, tob_deg=xd, K=k)
    if m == 0:
        return float("nan"), float("nan"), k
    seg = np.arange(m)[:, None] + np.arange(SEG)[None, :]
    cs = xc[:, seg].transpose(1, 0, 2)        # [M, 15, 30]
    ds = xd[:, seg].transpose(1, 0, 2)
    # STOI.py:129-139
    alpha = np.linalg.norm(cs, axis=2, keepdims=True) / (np.linalg.norm(ds, axis=2, keepdims=True) + 1e-9)
    dn = np.minimum(ds * alpha, cs * CLIP)
    stoi = np.sum(_normalize(cs, 2) * _normalize(dn, 2)) / N_BANDS / m
    ce = _normalize(_normalize(cs, 2), 1)     # STOI.py:178-181
    de = _normalize(_normalize(ds, 2), 1)
    estoi = np.sum(ce * de) / SEG / m
    return float(stoi), float(estoi), k


def stoi_item(clean: np.ndarray, deg: np.ndarray, sample_rate: int = 10000, taps: dict | None = None):
    """base.py:16-21 (resample iff sample_rate != 10000) then STOI.py:153-198."""
    c = resample(clean, sample_rate, FS)
    d = resample(deg, sample_rate, FS)
    return stoi_item_10k(c, d, taps)


def stoi_batch(clean: np.ndarray, deg: np.ndarray, sample_rate: int = 10000, lengths=None):
    clean = np.atleast_2d(clean)
    deg = np.atleast_2d(deg)
    b = clean.shape[0]
    out = np.empty((b, 2), np.float64)
    ks = np.empty(b, np.int64)
    for i in range(b):
        n = clean.shape[1] if lengths is None else int(lengths[i])
        s, e, k = stoi_item(clean[i, :n], deg[i, :n], sample_rate)
        out[i] = (s, e)
        ks[i] = k
    return out[:, 0], out[:, 1], ks
