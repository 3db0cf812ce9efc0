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
