import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The tests need the in-tree libfsem_b200.so (the product has no fallback).  It is git-ignored, so a fresh
    checkout does not have it: build it here once if it is missing or older than its sources and nvcc exists
    (cross-compiles for sm_100a without a GPU)."""
    from fast_speech_enhancement_metrics_b200 import build as b
    try:
        if b.is_stale():
            b.build()
    except Exception as exc:   # no nvcc: the tests that need the library report the missing .so themselves
        sys.stderr.write("libfsem_b200.so not built: %s\n" % exc)


@pytest.fixture(scope="session")
def golden_pesq():
    return dict(np.load(os.path.join(GOLDEN_DIR, "golden_pesq.npz")))


@pytest.fixture(scope="session")
def golden_stoi():
    return dict(np.load(os.path.join(GOLDEN_DIR, "golden_stoi.npz")))


def unpack_masks(golden, name):
    """Per-item boolean silent-frame masks of a golden STOI case."""
    bits = golden[name + "/mask_bits"]
    nbytes = golden[name + "/mask_nbytes"]
    out, off = [], 0
    for nb in nbytes:
        out.append(np.unpackbits(bits[off:off + nb]).astype(bool))
        off += nb
    return out
