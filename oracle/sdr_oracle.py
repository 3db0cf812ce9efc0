"""Float64 numpy restatement of the reference's SDR (fast_se_metrics/SDR.py:7-97).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Correlations through zero-padded FFTs as the reference does
(SDR.py:34-49), the symmetric Toeplitz system with scipy's Levinson solver (the reference uses a float32 Cholesky,
SDR.py:7-31); the reference's float32 arithmetic leaves it ~5e-3 dB (at 20 dB) from this float64 value."""
import numpy as np
from scipy.linalg import solve_toeplitz

FILTER_LENGTH = 512


def sdr_item(clean: np.ndarray, deg: np.ndarray) -> float:
    c = np.asarray(clean, np.float64)
    d = np.asarray(deg, np.float64)
    c = c / max(np.linalg.norm(c), 1e-6)                       # SDR.py:65-71
    d = d / max(np.linalg.norm(d), 1e-6)
    n = c.shape[0]
    nfft = 1 << int(np.ceil(np.log2(2 * n - 1)))
    cf, df = np.fft.rfft(c, nfft), np.fft.rfft(d, nfft)
    r0 = np.fft.irfft(np.abs(cf) ** 2, nfft)[:FILTER_LENGTH]
    b = np.fft.irfft(np.conj(cf) * df, nfft)[:FILTER_LENGTH]
    sol = solve_toeplitz(r0, b)
    coh = float(b @ sol)                                        # SDR.py:88
    ratio = coh / max(1.0 - coh, 1e-8)
    return float(10.0 * np.log10(max(ratio, 1e-8)))            # SDR.py:91-95


def sdr_batch(clean: np.ndarray, deg: np.ndarray, lengths=None, sample_rate: int = 16000) -> np.ndarray:
    """sample_rate != 16000: resample-on-ingest first (base.py:19-20, torchaudio sinc-Hann kernel)."""
    from .stoi_oracle import resample
    clean, deg = np.atleast_2d(clean), np.atleast_2d(deg)
    out = np.empty(clean.shape[0])
    for i in range(clean.shape[0]):
        n = clean.shape[1] if lengths is None else int(lengths[i])
        out[i] = sdr_item(resample(clean[i, :n], sample_rate, 16000), resample(deg[i, :n], sample_rate, 16000))
    return out
