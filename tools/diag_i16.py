#!/usr/bin/env python
"""Where does the int16 end-to-end time go?  Device-resident int16 scoring at chunk-sized and full batches, and the host
entry point at two batch sizes (python tools/diag_i16.py)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fast_speech_enhancement_metrics_b200 import PESQ, STOI, _lib, score_pesq_stoi, score_pesq_stoi_tensors
from tools.config_sweep import make

pesq, stoi = PESQ(16000, True), STOI(16000, True)
n = 160000
for dtype in (torch.float32, torch.int16):
    for b in (100, 200, 400, 8192):
        c, d = make(b, n, 11)
        if dtype == torch.int16:
            s = 32767.0 / float(torch.maximum(c.abs().max(), d.abs().max()))
            c, d = (c * s).round().to(torch.int16), (d * s).round().to(torch.int16)
        for _ in range(3):
            score_pesq_stoi_tensors(pesq, stoi, c, d)
        torch.cuda.synchronize()
        _lib.profile_reset(); _lib.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            score_pesq_stoi_tensors(pesq, stoi, c, d)
        e1.record(); torch.cuda.synchronize()
        _lib.profile_enable(False)
        prof = {k: round(v[0] / reps, 3) for k, v in _lib.profile_read().items() if v[1]}
        print(str(dtype), b, "ms %.3f" % (e0.elapsed_time(e1) / reps), prof, flush=True)
        if b == 8192 or b == 4096:
            for bb in (8192, 4096):
                hc = torch.empty((bb, n), dtype=dtype, pin_memory=True).copy_(c[:bb])
                hd = torch.empty((bb, n), dtype=dtype, pin_memory=True).copy_(d[:bb])
                score_pesq_stoi(pesq, stoi, hc, hd)
                t0 = time.perf_counter()
                for _ in range(3):
                    score_pesq_stoi(pesq, stoi, hc, hd)
                dt = (time.perf_counter() - t0) / 3
                gb = 2 * bb * n * hc.element_size() / 1e9
                print("   host entry", str(dtype), bb, "ms %.1f  %.1f GB/s  %.0f audio-s/s" % (dt * 1e3, gb / dt, bb * 10 / dt), flush=True)
                del hc, hd
        del c, d
        torch.cuda.empty_cache()
