"""Smallest end-to-end invocation for compute-sanitizer (memcheck): PESQ + STOI, fixed and ragged lengths."""
import sys; sys.path.insert(0, ".")
import torch
from fast_speech_enhancement_metrics_b200 import PESQ, STOI, score_pesq_stoi
from fast_speech_enhancement_metrics_b200.synth import synth_batch
c, d, _ = synth_batch(5, 3, 24000)
c, d = torch.from_numpy(c).cuda(), torch.from_numpy(d).cuda()
p, s = PESQ(16000, True), STOI(16000, True)
print(p(c, d)); print(s(c, d))
print(p(c, d, lengths=[24000, 5376, 17001])); print(s(c, d, lengths=[24000, 9000, 17001]))
print(score_pesq_stoi(p, s, c.cpu(), d.cpu()))
s10 = STOI(10000, True)
print(s10(c[:, :20000].contiguous(), d[:, :20000].contiguous()))
p8 = PESQ(8000, True)
print(p8(c[:, :12000].contiguous(), d[:, :12000].contiguous()))
torch.cuda.synchronize()
print("done")
# round 2: the graph path (bulk-copied Bark tiles, sliced branches), a 33-item batch (tiled IIR kernel), LSD / SDR
from fast_speech_enhancement_metrics_b200 import LSD, SDR, CapturedScorer
c33, d33, _ = synth_batch(6, 33, 20004)
c33, d33 = torch.from_numpy(c33).cuda(), torch.from_numpy(d33).cuda()
sc = CapturedScorer(p, s, c33, d33, slices=3)
print(sc()[:2]); sc.close()
sc = CapturedScorer(p, s, c, d, lengths=[24000, 5400, 17001])
print(sc()); sc.close()
print(LSD(16000, True)(c, d)); print(SDR(16000, True)(c, d))
torch.cuda.synchronize()
print("done round 2")
