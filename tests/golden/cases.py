"""Seeded input cases shared by make_golden.py (reference outputs) and the tests.

Every case is a pure function of numpy's PCG64 stream, so the fixtures store only
the reference's outputs.  A case is (clean[B, n] f32, degraded[B, n] f32, lengths or
None[, sample_rate]); `lengths` means "item i is row i cut to lengths[i]" (the
reference was called per item on the slice).
"""
from __future__ import annotations

import numpy as np
from scipy.signal import lfilter

from fast_speech_enhancement_metrics_b200.synth import synth_batch


def _white(seed, b, n, scale=1.0):
    rng = np.random.default_rng(seed)
    return (scale * rng.standard_normal((b, n))).astype(np.float32)


def pesq_cases():
    cases = {}
    c, d, _ = synth_batch(101, 8, 32000)
    cases["speech2s"] = (c, d, None)
    c, d, _ = synth_batch(102, 4, 160000)
    cases["speech10s"] = (c, d, None)                       # README shape (BASELINE configs[0])
    c, d, _ = synth_batch(103, 6, 40001)
    cases["ragged"] = (c, d, [40001, 20037, 5376, 5377, 5631, 33333])   # incl. the minimum length (T = 20)
    c, d, _ = synth_batch(104, 2, 24000)
    cases["identical"] = (c, c.copy(), None)                # reference gives 4.6438887...
    cases["white"] = (_white(105, 3, 48000), _white(106, 3, 48000), None)
    cw = _white(107, 3, 48000)
    cases["white_plus_noise"] = (cw, (cw + 0.3 * _white(108, 3, 48000)).astype(np.float32), None)
    # low-frequency-heavy + DC offset: stresses the band-pass stop band
    rng = np.random.default_rng(109)
    lf = lfilter([1.0], [1.0, -0.999], rng.standard_normal((2, 48000)), axis=1) * 0.01 + 0.5
    lf = lf.astype(np.float32)
    cases["lf_dc"] = (lf, (lf + 0.05 * _white(110, 2, 48000)).astype(np.float32), None)
    # tiny and huge amplitudes (level alignment must cancel them)
    c, d, _ = synth_batch(111, 2, 32000)
    cases["scaled"] = (np.stack([c[0] * 1e-4, c[1] * 3e3]).astype(np.float32),
                       np.stack([d[0] * 1e-4, d[1] * 3e3]).astype(np.float32), None)
    z = np.zeros((1, 16000), np.float32)
    cases["all_zero"] = (z, z.copy(), None)                 # nan
    cases["deg_zero"] = (_white(112, 1, 16000), z.copy(), None)   # nan
    return cases


def pesq_rate_cases():
    """PESQ with resample-on-ingest (base.py:13,19-20): (clean, degraded, lengths, sample_rate)."""
    cases = {}
    c, d, _ = synth_batch(121, 3, 16000, fs=8000)
    cases["speech8k_2s"] = (c, d, None, 8000)               # 1:2 up-sampling
    c, d, _ = synth_batch(122, 2, 96000, fs=48000)
    cases["speech48k_2s"] = (c, d, None, 48000)             # 3:1 down-sampling
    c, d, _ = synth_batch(123, 2, 44100, fs=44100)
    cases["speech44k1_1s"] = (c, d, None, 44100)            # 441:160
    c, d, _ = synth_batch(124, 3, 48000, fs=24000)
    cases["ragged24k"] = (c, d, [48000, 30001, 12345], 24000)
    return cases


def lsd_cases():
    """LSD (fast_se_metrics/LSD.py): (clean, degraded, lengths)."""
    cases = {}
    c, d, _ = synth_batch(301, 6, 32000)
    cases["speech2s"] = (c, d, None)
    c, d, _ = synth_batch(302, 4, 40001)
    cases["ragged"] = (c, d, [40001, 20037, 511, 33333])
    cw = _white(303, 2, 16000)
    cases["white_scaled"] = (cw, (0.1 * cw + 0.02 * _white(304, 2, 16000)).astype(np.float32), None)
    return cases


def lsd_rate_cases():
    """LSD with resample-on-ingest to 16 kHz (base.py:13,19-20): (clean, degraded, lengths, sample_rate)."""
    cases = {}
    c, d, _ = synth_batch(311, 3, 16000, fs=8000)
    cases["speech8k_2s"] = (c, d, None, 8000)
    c, d, _ = synth_batch(312, 3, 96000, fs=48000)
    cases["ragged48k"] = (c, d, [96000, 60001, 24000], 48000)
    return cases


def sdr_rate_cases():
    """SDR with resample-on-ingest to 16 kHz (base.py:13,19-20): (clean, degraded, lengths, sample_rate).
    Down-sampling rates only: an up-sampled (band-limited) signal makes the 512 x 512 autocorrelation matrix
    numerically singular -- the reference's own float32 Cholesky fails on 8 kHz input and its fallback
    `torch.linalg.solve` dies inside MKL (SLASWP parameter error), so there is no reference value to pin."""
    cases = {}
    c, d, _ = synth_batch(411, 3, 64000, fs=32000)
    cases["speech32k_2s"] = (c, d, None, 32000)
    c, d, _ = synth_batch(412, 3, 96000, fs=48000)
    cases["ragged48k"] = (c, d, [96000, 60001, 24000], 48000)
    return cases


def sdr_cases():
    """SDR (fast_se_metrics/SDR.py): (clean, degraded, lengths)."""
    cases = {}
    c, d, _ = synth_batch(401, 6, 32000)
    cases["speech2s"] = (c, d, None)
    c, d, _ = synth_batch(402, 3, 160000)
    cases["speech10s"] = (c, d, None)
    c, d, _ = synth_batch(403, 4, 40001)
    cases["ragged"] = (c, d, [40001, 20037, 2049, 33333])
    cw = _white(404, 2, 16000)
    cases["white_scaled"] = (cw, (0.1 * cw + 0.02 * _white(405, 2, 16000)).astype(np.float32), None)
    return cases


def stoi_cases():
    cases = {}
    c, d, _ = synth_batch(201, 8, 30000, fs=10000)
    cases["speech10k_3s"] = (c, d, None, 10000)             # no resampler: mask must be bit-exact
    c, d, _ = synth_batch(202, 8, 48000)
    cases["speech16k_3s"] = (c, d, None, 16000)
    c, d, _ = synth_batch(203, 4, 160000)
    cases["speech16k_10s"] = (c, d, None, 16000)            # README shape
    c, d, _ = synth_batch(204, 6, 64000)
    cases["ragged16k"] = (c, d, [64000, 47001, 16000, 30011, 8000, 12345], 16000)
    c, d, _ = synth_batch(205, 5, 40000, fs=10000)
    cases["ragged10k"] = (c, d, [40000, 4224, 4223, 20001, 9999], 10000)   # 4224 -> exactly 32 frames
    cw = _white(206, 3, 40000)
    cases["white16k"] = (cw, (cw + 0.5 * _white(207, 3, 40000)).astype(np.float32), None, 16000)
    cases["white_uncorr10k"] = (_white(208, 2, 30000), _white(209, 2, 30000), None, 10000)
    c, d, _ = synth_batch(210, 2, 30000, fs=10000)
    cases["identical10k"] = (c, c.copy(), None, 10000)      # 1.0
    # an item with too few kept frames inside a batch that has segments -> nan for that item
    c, d, _ = synth_batch(211, 3, 30000, fs=10000)
    c[1, 3000:] *= 1e-5
    d[1, 3000:] *= 1e-5
    cases["one_short10k"] = (c, d, None, 10000)
    c, d, _ = synth_batch(212, 2, 44100, fs=44100)
    cases["speech8k"] = (c[:, :24000], d[:, :24000], None, 8000)   # up-sampling 8k -> 10k (4:5)
    return cases
