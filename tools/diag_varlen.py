import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from fast_speech_enhancement_metrics_b200 import PESQ
from fast_speech_enhancement_metrics_b200.synth import synth_item
from oracle import pesq_oracle as po
rng = np.random.default_rng(4321)
lens = [int(x) for x in rng.integers(16000, 480001, size=24)]
lens[0], lens[1] = 480000, 16000
nmax = max(lens)
clean = np.zeros((len(lens), nmax), np.float32); deg = np.zeros_like(clean)
for i, n in enumerate(lens):
    clean[i, :n], deg[i, :n], _ = synth_item(rng, n)
pesq = PESQ(16000, True)
want = po.pesq_batch(clean, deg, lens)
c, d = torch.from_numpy(clean).cuda(), torch.from_numpy(deg).cuda()
got0 = np.array([r["PESQ"] for r in pesq(c, d, lengths=lens)])
clean2 = clean.copy(); deg2 = deg.copy()
for i, n in enumerate(lens):
    clean2[i, n:] = 7.0; deg2[i, n:] = -3.0
got1 = np.array([r["PESQ"] for r in pesq(torch.from_numpy(clean2).cuda(), torch.from_numpy(deg2).cuda(), lengths=lens)])
sl = np.array([pesq(c[i:i+1,:n].contiguous(), d[i:i+1,:n].contiguous())[0]["PESQ"] for i,n in enumerate(lens)])
for i,n in enumerate(lens):
    print(i, n, n%256, '%.6f'%want[i], 'zero-pad %.2e'%(got0[i]-want[i]), 'garbage %.2e'%(got1[i]-want[i]), 'sliced %.2e'%(sl[i]-want[i]))
