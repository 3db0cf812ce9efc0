"""In-tree nvcc build of libfsem_b200.so (sm_100a only).

    python -m fast_speech_enhancement_metrics_b200.build [--verbose]

The library is built next to this file so that it travels with the repo snapshot
to the GPU box; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_NAME = "libfsem_b200.so"
LIB_PATH = os.path.join(HERE, LIB_NAME)
SOURCES = ["fsem_api.cu"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith(".cuh")) + [os.path.join("..", "..", "include", "fsem.h")]


def find_nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found (need CUDA 12.9 for sm_100a)")
    return cand


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=()) -> str:
    if not force and not is_stale():
        return LIB_PATH
    cmd = [
        find_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
        "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
        "-Xptxas", "-v" if verbose else "-warn-spills",
        "-o", LIB_PATH,
    ] + ["-D" + d for d in defines] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building %s" % LIB_NAME)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose="--verbose" in sys.argv, defines=[a[2:] for a in sys.argv[1:] if a.startswith("-D")]))
