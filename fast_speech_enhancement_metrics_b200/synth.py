"""Seeded synthetic speech-like audio (numpy PCG64 -- stable across numpy versions).

Used by the tests, the golden-fixture generator and bench.py.  Recipe (SURVEY.md
section 8d): `clean` = white Gaussian noise through a one-pole low-pass (pole 0.95)
times a syllabic envelope max(0, sin(2*pi*f*t + phi))^2 with f ~ U[1.5, 4] Hz, peak
~ U[0.1, 0.3], plus a -60 dB white floor so pauses are quiet but not exactly zero
(silent-frame compaction is exercised); `degraded` = clean + white noise at a
per-item SNR ~ U[snr_lo, snr_hi] dB (the reference benchmark's range,
benchmarking/dataloading.py:63-72).
"""
from __future__ import annotations

import numpy as np
from scipy.signal import lfilter


def synth_item(rng: np.random.Generator, n: int, fs: int = 16000, snr_db: float | None = None,
               snr_range=(-5.0, 25.0)):
    white = rng.standard_normal(n)
    lp = lfilter([1.0], [1.0, -0.95], white)
    f = rng.uniform(1.5, 4.0)
    phi = rng.uniform(0.0, 2.0 * np.pi)
    t = np.arange(n) / fs
    env = np.maximum(0.0, np.sin(2.0 * np.pi * f * t + phi)) ** 2
    clean = lp * env
    peak = rng.uniform(0.1, 0.3)
    clean = clean * (peak / max(np.abs(clean).max(), 1e-12))
    clean = clean + (peak * 1e-3) * rng.standard_normal(n)
    if snr_db is None:
        snr_db = rng.uniform(*snr_range)
    noise = rng.standard_normal(n)
    p_clean = np.mean(clean ** 2)
    p_noise = np.mean(noise ** 2)
    noise = noise * np.sqrt(p_clean / (p_noise * 10.0 ** (snr_db / 10.0)))
    return clean.astype(np.float32), (clean + noise).astype(np.float32), float(snr_db)


def synth_batch(seed: int, batch: int, n: int, fs: int = 16000, snr_range=(-5.0, 25.0)):
    """Returns (clean[batch, n] f32, degraded[batch, n] f32, snr_db[batch])."""
    rng = np.random.default_rng(seed)
    clean = np.empty((batch, n), np.float32)
    deg = np.empty((batch, n), np.float32)
    snr = np.empty(batch, np.float64)
    for i in range(batch):
        clean[i], deg[i], snr[i] = synth_item(rng, n, fs, snr_range=snr_range)
    return clean, deg, snr
