#!/usr/bin/env python
"""Recipe for oracle/_ref: the UNMODIFIED reference implementation of the hot path, staged so that it can
run next to the GPU (test / bench infrastructure, never part of the product).

    python oracle/make_ref.py            # build container only: needs /root/reference

The reference is pure Python (0 lines of C/C++/CUDA), so "building" it is copying the module files of the
path verbatim from where they lie under /root/reference into the git-ignored oracle/_ref/fast_se_metrics/
(outputs only there; nothing is copied into tracked files):

    base.py PESQ.py STOI.py LSD.py SDR.py utils/bark.py utils/loudness.py

plus a one-line package `__init__` that does NOT import DNSMOS / SpeechBERTScore: the reference's own
`fast_se_metrics/__init__.py:5-6` pulls in `transformers` (14 s import) and flips
`torch.set_float32_matmul_precision("high")` for metrics that are outside the path.  The arithmetic files are
byte-identical to the reference's (checked below by hash), so `oracle/_ref` IS the reference's CPU path:
`fast_se_metrics.PESQ.PESQ(16000, use_gpu=False)` / `fast_se_metrics.STOI.STOI(16000, use_gpu=False)`.

oracle/_ref/ is listed in .gitignore (stays out of history) but not in .gpurunignore (travels to the GPU
box like the built .so).  `bench.py --impl reference` and `bench.py`'s `cpu_baseline` leg time it
(kind = "reference"); without it they fall back to the numpy port (kind = "port").
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/fast_se_metrics"
REF_DST = os.path.join(HERE, "_ref")
FILES = ["base.py", "PESQ.py", "STOI.py", "LSD.py", "SDR.py", os.path.join("utils", "bark.py"),
         os.path.join("utils", "loudness.py")]
INIT = ('"""oracle/_ref: verbatim copies of the reference\'s hot-path modules (see oracle/make_ref.py); this __init__ '
        'replaces fast_se_metrics/__init__.py so that DNSMOS / SpeechBERTScore (transformers) are not imported."""\n')


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def available() -> bool:
    """True when oracle/_ref holds every module of the path."""
    pkg = os.path.join(REF_DST, "fast_se_metrics")
    return all(os.path.exists(os.path.join(pkg, f)) for f in FILES + ["__init__.py"])


def build(verbose: bool = False) -> bool:
    """Stage the reference modules; returns False (and leaves any existing oracle/_ref alone) when
    /root/reference is absent, e.g. on the GPU box, where the prebuilt copy is used."""
    if not os.path.isdir(REF_SRC):
        return available()
    pkg = os.path.join(REF_DST, "fast_se_metrics")
    os.makedirs(os.path.join(pkg, "utils"), exist_ok=True)
    manifest = []
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(pkg, rel)
        shutil.copyfile(src, dst)
        if _sha(src) != _sha(dst):
            raise RuntimeError("oracle/_ref: %s differs from the reference after copying" % rel)
        manifest.append("%s  %s" % (_sha(dst), rel))
    with open(os.path.join(pkg, "__init__.py"), "w") as f:
        f.write(INIT)
    with open(os.path.join(pkg, "utils", "__init__.py"), "w") as f:
        f.write("")
    with open(os.path.join(REF_DST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print("oracle/_ref staged from %s:\n%s" % (REF_SRC, "\n".join(manifest)))
    return True


def load():
    """(PESQ, STOI) classes of the staged reference; raises ImportError when oracle/_ref is missing."""
    if not available():
        raise ImportError("oracle/_ref is not staged (python oracle/make_ref.py in the build container)")
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    from fast_se_metrics.PESQ import PESQ   # noqa: E402
    from fast_se_metrics.STOI import STOI   # noqa: E402
    return PESQ, STOI


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref available:", ok)
