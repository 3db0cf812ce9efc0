#!/usr/bin/env python
"""Per-item PESQ deviations of the bench's sample items on the 8 shards of an 8-GPU run, reproduced on one GPU:
    python tools/diag_parity_shards.py [--items 32]
For every shard seed 1000 + r (1024 x 10 s): CUDA vs float64 oracle vs the staged reference (oracle/_ref), worst items listed."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fast_speech_enhancement_metrics_b200 import PESQ  # noqa: E402
from oracle import make_ref, pesq_oracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=32)
    ap.add_argument("--shards", type=int, default=8)
    ap.add_argument("--local-batch", type=int, default=1024)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    pesq = PESQ(16000, use_gpu=True)
    ref = make_ref.load()[0](16000, use_gpu=False) if make_ref.available() else None
    for r in range(args.shards):
        clean, deg = bench.make_shard(args.local_batch, 160000, 1000 + r, dev)
        mos, _ = pesq.score_tensors(clean, deg)
        idx = np.unique(np.linspace(0, args.local_batch - 1, args.items).round().astype(np.int64))
        t = torch.from_numpy(idx).to(dev)
        c, d = clean[t].cpu(), deg[t].cpu()
        got = mos[t].double().cpu().numpy()
        orc = pesq_oracle.pesq_batch(c.numpy(), d.numpy())
        # the same items alone in a small batch (different IIR chunk grid)
        small = pesq.score_tensors(clean[t].contiguous(), deg[t].contiguous())[0].double().cpu().numpy()
        line = "shard %d: max|cuda-oracle| %.2e  max|cuda(batch %d)-cuda(batch %d)| %.2e" % (
            r, np.max(np.abs(got - orc)), args.local_batch, len(idx), np.max(np.abs(got - small)))
        if ref is not None:
            rv = np.array([x["PESQ"] for x in ref(c, d)])
            line += "  max|cuda-ref| %.2e  max|ref-oracle| %.2e" % (np.max(np.abs(got - rv)), np.max(np.abs(rv - orc)))
            w = int(np.argmax(np.abs(got - rv)))
            line += "  worst item %d: cuda %.6f oracle %.6f ref %.6f small-batch cuda %.6f" % (idx[w], got[w], orc[w], rv[w], small[w])
        print(line, flush=True)
        del clean, deg


if __name__ == "__main__":
    main()
