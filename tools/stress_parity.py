"""One-off stress: random batches (with lengths) where each FFT warp walks across several items; all items vs oracle."""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
from fast_speech_enhancement_metrics_b200 import PESQ, STOI
from fast_speech_enhancement_metrics_b200.synth import synth_item
from oracle import pesq_oracle as po, stoi_oracle as so
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 7)
pesq, stoi = PESQ(16000, True), STOI(16000, True)
worst = [0, 0, 0]; bad = 0
for trial in range(12):
    b = int(rng.integers(300, 900)); n = int(rng.integers(12000, 40000))
    clean = np.zeros((b, n), np.float32); deg = np.zeros_like(clean)
    base_c, base_d, _ = synth_item(rng, n * 4)
    for i in range(b):
        o = int(rng.integers(0, 3 * n)); g = rng.uniform(0.3, 2.0)
        clean[i] = base_c[o:o + n] * g; deg[i] = base_d[o:o + n] * g
    lens = [int(x) for x in rng.integers(9000, n + 1, size=b)] if trial % 2 else None
    c, d = torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda()
    gp = np.array([r["PESQ"] for r in pesq(c, d, lengths=lens)])
    rs = stoi(c, d, lengths=lens)
    pick = rng.choice(b, size=60, replace=False)
    wl = None if lens is None else [lens[i] for i in pick]
    wp = po.pesq_batch(clean[pick], deg[pick], wl)
    ws, we, wk = so.stoi_batch(clean[pick], deg[pick], 16000, wl)
    dp = np.abs(gp[pick] - wp); ds = np.abs(np.array([rs[i]["STOI"] for i in pick]) - ws); de = np.abs(np.array([rs[i]["ESTOI"] for i in pick]) - we)
    ok = ~np.isnan(ws)
    worst = [max(worst[0], dp.max()), max(worst[1], np.nanmax(ds[ok], initial=0)), max(worst[2], np.nanmax(de[ok], initial=0))]
    bad += int(not np.array_equal(stoi.last_kept_frames.numpy()[pick], wk)) + int(np.any(np.isnan(ws) != np.isnan(np.array([rs[i]["STOI"] for i in pick]))))
    print(trial, b, n, lens is not None, "%.2e %.2e %.2e" % (dp.max(), ds[ok].max() if ok.any() else 0, de[ok].max() if ok.any() else 0), flush=True)
print("worst", worst, "mismatches", bad)
