import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from fast_speech_enhancement_metrics_b200 import SDR
from fast_speech_enhancement_metrics_b200.synth import synth_batch
from oracle import sdr_oracle
c, d, _ = synth_batch(402, 3, 160000)
m = SDR(16000, use_gpu=True)
tc, td = torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda()
outs = [m.score_tensors(tc, td).cpu().numpy() for _ in range(4)]
print("mode", os.environ.get("FSEM_SDR_SIMT"), os.environ.get("FSEM_SDR_NO_TMA"))
for o in outs: print(o)
print("oracle", sdr_oracle.sdr_batch(c, d))
c2, d2 = tc.clone(), td.clone()
print("clone ", m.score_tensors(c2, d2).cpu().numpy())
big_c = torch.zeros(5, 160000 + 64, device="cuda"); big_d = torch.zeros_like(big_c)
big_c[1:4, 8:160008] = tc; big_d[1:4, 8:160008] = td
print("view  ", m.score_tensors(big_c[1:4, 8:160008], big_d[1:4, 8:160008]).cpu().numpy())
