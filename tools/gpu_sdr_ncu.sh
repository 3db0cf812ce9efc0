#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/sdr_one.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from fast_speech_enhancement_metrics_b200 import SDR
import bench
clean, deg = bench.make_shard(1024, 160000, 1000, torch.device("cuda", 0))
m = SDR(16000, use_gpu=True)
for _ in range(3): out = m.score_tensors(clean, deg)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sdr_corr_tc" -s 2 -c 1 -o gpurun_out/sdrtc_full -f python /tmp/sdr_one.py > gpurun_out/sdrtc_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/sdrtc_full.ncu-rep --page raw --csv > gpurun_out/sdrtc_full_raw.csv 2>/dev/null
