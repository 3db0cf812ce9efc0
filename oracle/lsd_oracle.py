"""Float64 numpy restatement of the reference's LSD (fast_se_metrics/LSD.py:18-52).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  torch.stft(center=True, pad_mode="constant", n_fft=512,
hop=256, hann periodic): the signal is zero-padded by 256 samples on both sides; T = 1 + n // 256 frames."""
import numpy as np

EPS = 1e-8
_HANN = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(512) / 512)


def lsd_item(clean: np.ndarray, deg: np.ndarray) -> float:
    c = np.asarray(clean, np.float64)
    d = np.asarray(deg, np.float64)
    alpha = np.float32(np.dot(c, d)) / (np.float32(np.dot(d, d)) + np.float32(EPS))      # LSD.py:37-39 (float32 ratio)
    d = d * np.float64(alpha)
    n = c.shape[0]
    t = 1 + n // 256
    idx = np.arange(512)[None, :] + 256 * np.arange(t)[:, None]
    spec = []
    for x in (c, d):
        xp = np.concatenate([np.zeros(256), x, np.zeros(256)])
        spec.append(np.abs(np.fft.rfft(xp[idx] * _HANN[None, :], axis=1)))            # [T, 257]
    lsd = np.log(spec[0] ** 2 / (spec[1] + EPS) ** 2 + EPS) ** 2                       # LSD.py:47-49
    return float(np.mean(np.sqrt(np.mean(lsd, axis=1))))                              # LSD.py:50


def lsd_batch(clean: np.ndarray, deg: np.ndarray, lengths=None, sample_rate: int = 16000) -> np.ndarray:
    """sample_rate != 16000: resample-on-ingest first (base.py:19-20, torchaudio sinc-Hann kernel)."""
    from .stoi_oracle import resample
    clean, deg = np.atleast_2d(clean), np.atleast_2d(deg)
    out = np.empty(clean.shape[0])
    for i in range(clean.shape[0]):
        n = clean.shape[1] if lengths is None else int(lengths[i])
        out[i] = lsd_item(resample(clean[i, :n], sample_rate, 16000), resample(deg[i, :n], sample_rate, 16000))
    return out
