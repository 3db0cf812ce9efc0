"""LSD: oracle vs the reference's outputs (CPU) and CUDA vs both (GPU)."""
import os

import numpy as np
import pytest

from oracle import lsd_oracle as lo
from tests.conftest import GOLDEN_DIR
from tests.golden.cases import lsd_cases, lsd_rate_cases

CASES = lsd_cases()
RATE_CASES = lsd_rate_cases()


@pytest.fixture(scope="module")
def golden_lsd():
    return dict(np.load(os.path.join(GOLDEN_DIR, "golden_lsd.npz")))


@pytest.mark.parametrize("name", sorted(CASES))
def test_lsd_oracle_matches_reference(name, golden_lsd):
    clean, deg, lengths = CASES[name]
    got = lo.lsd_batch(clean, deg, lengths)
    assert np.max(np.abs(got - golden_lsd[name])) <= 2e-4 * np.max(golden_lsd[name])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_lsd_matches_reference_and_oracle(name, golden_lsd):
    import torch

    from fast_speech_enhancement_metrics_b200 import LSD
    clean, deg, lengths = CASES[name]
    metric = LSD(16000, use_gpu=True)
    got = np.array([r["LSD"] for r in metric(torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda(), lengths=lengths)])
    host = np.array([r["LSD"] for r in metric(torch.from_numpy(clean), torch.from_numpy(deg), lengths=lengths)])
    want = golden_lsd[name]
    # LSD averages log-ratios over bins whose degraded magnitude can be ~1e-8 (the eps floor): float32 FFT round-off in
    # those bins moves the score at the 1e-4 relative level in the reference's own float32 path as well
    assert np.max(np.abs(got - want) / want) <= 5e-4
    assert np.max(np.abs(got - lo.lsd_batch(clean, deg, lengths)) / want) <= 5e-4
    assert np.array_equal(got, host)


@pytest.mark.parametrize("name", sorted(RATE_CASES))
def test_lsd_oracle_resample_on_ingest_matches_reference(name, golden_lsd):
    clean, deg, lengths, fs = RATE_CASES[name]
    got = lo.lsd_batch(clean, deg, lengths, sample_rate=fs)
    want = golden_lsd["rate/" + name]
    assert np.max(np.abs(got - want) / want) <= 5e-4


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(RATE_CASES))
def test_lsd_resample_on_ingest(name, golden_lsd):
    """LSD(sample_rate != 16000): resample-on-ingest through the library's polyphase kernel (base.py:13,19-20)."""
    import torch

    from fast_speech_enhancement_metrics_b200 import LSD
    clean, deg, lengths, fs = RATE_CASES[name]
    metric = LSD(fs, use_gpu=True)
    got = np.array([r["LSD"] for r in metric(torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda(), lengths=lengths)])
    host = np.array([r["LSD"] for r in metric(torch.from_numpy(clean), torch.from_numpy(deg), lengths=lengths)])
    want = golden_lsd["rate/" + name]
    assert np.max(np.abs(got - want) / want) <= 5e-4
    assert np.max(np.abs(got - lo.lsd_batch(clean, deg, lengths, sample_rate=fs)) / want) <= 5e-4
    assert np.array_equal(got, host)
