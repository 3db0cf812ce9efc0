"""SDR: oracle vs the reference's outputs (CPU) and CUDA vs both (GPU).

Tolerance: the reference solves an ill-conditioned 512 x 512 Toeplitz system in float32 (its own CPU<->GPU test allows
0.1 dB, tests/test_cuda.py:21); the float64 oracle is within 2e-2 dB of it on these cases, the CUDA path (fp64 solve) is
held to 5e-2 dB of the reference and 2e-2 dB of the oracle."""
import os

import numpy as np
import pytest

from oracle import sdr_oracle as sd
from tests.conftest import GOLDEN_DIR
from tests.golden.cases import sdr_cases, sdr_rate_cases

CASES = sdr_cases()
RATE_CASES = sdr_rate_cases()


@pytest.fixture(scope="module")
def golden_sdr():
    return dict(np.load(os.path.join(GOLDEN_DIR, "golden_sdr.npz")))


@pytest.mark.parametrize("name", sorted(CASES))
def test_sdr_oracle_matches_reference(name, golden_sdr):
    clean, deg, lengths = CASES[name]
    got = sd.sdr_batch(clean, deg, lengths)
    assert np.max(np.abs(got - golden_sdr[name])) <= 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_sdr_matches_reference_and_oracle(name, golden_sdr):
    import torch

    from fast_speech_enhancement_metrics_b200 import SDR
    clean, deg, lengths = CASES[name]
    metric = SDR(16000, use_gpu=True)
    got = np.array([r["SDR"] for r in metric(torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda(), lengths=lengths)])
    host = np.array([r["SDR"] for r in metric(torch.from_numpy(clean), torch.from_numpy(deg), lengths=lengths)])
    assert np.max(np.abs(got - golden_sdr[name])) <= 5e-2
    assert np.max(np.abs(got - sd.sdr_batch(clean, deg, lengths))) <= 2e-2
    assert np.array_equal(got, host)


@pytest.mark.parametrize("name", sorted(RATE_CASES))
def test_sdr_oracle_resample_on_ingest_matches_reference(name, golden_sdr):
    clean, deg, lengths, fs = RATE_CASES[name]
    got = sd.sdr_batch(clean, deg, lengths, sample_rate=fs)
    assert np.max(np.abs(got - golden_sdr["rate/" + name])) <= 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(RATE_CASES))
def test_sdr_resample_on_ingest(name, golden_sdr):
    """SDR(sample_rate != 16000): resample-on-ingest through the library's polyphase kernel (base.py:13,19-20)."""
    import torch

    from fast_speech_enhancement_metrics_b200 import SDR
    clean, deg, lengths, fs = RATE_CASES[name]
    metric = SDR(fs, use_gpu=True)
    got = np.array([r["SDR"] for r in metric(torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda(), lengths=lengths)])
    host = np.array([r["SDR"] for r in metric(torch.from_numpy(clean), torch.from_numpy(deg), lengths=lengths)])
    assert np.max(np.abs(got - golden_sdr["rate/" + name])) <= 5e-2
    assert np.max(np.abs(got - sd.sdr_batch(clean, deg, lengths, sample_rate=fs))) <= 2e-2
    assert np.array_equal(got, host)
