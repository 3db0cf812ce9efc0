#!/bin/bash
# SDR tensor-core kernel: correctness (tests) and A/B timing against the SIMT kernel at 8192 x 10 s
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sdr.py tests/test_gpu_parity.py -x -q -m gpu -k "sdr or int16_and_fp16" > gpurun_out/sdr_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/sdr_pytest.log
cat > /tmp/sdr_ab.py <<'PY'
import sys, os, json, torch
sys.path.insert(0, os.getcwd())
from fast_speech_enhancement_metrics_b200 import SDR, _lib
import bench
dev = torch.device("cuda", 0)
b = int(sys.argv[1])
clean, deg = bench.make_shard(b, 160000, 1000, dev)
m = SDR(16000, use_gpu=True)
for _ in range(2): out = m.score_tensors(clean, deg)
torch.cuda.synchronize()
_lib.profile_reset(); _lib.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): out = m.score_tensors(clean, deg)
e1.record(); torch.cuda.synchronize()
_lib.profile_enable(False)
p = _lib.profile_read()
print(json.dumps({"simt": os.environ.get("FSEM_SDR_SIMT", "0"), "batch": b, "ms_per_call": e0.elapsed_time(e1) / 3,
                  "corr_ms": p["sdr_corr_kernel"][0] / 3, "solve_ms": p["sdr_solve_kernel"][0] / 3,
                  "sdr_head": out[:4].tolist(), "finite": bool(torch.isfinite(out).all())}))
PY
timeout 300 python /tmp/sdr_ab.py 2048 > gpurun_out/sdr_tc.json 2> gpurun_out/sdr_tc.err; echo "tc rc=$?"; cat gpurun_out/sdr_tc.json; tail -3 gpurun_out/sdr_tc.err
FSEM_SDR_SIMT=1 timeout 300 python /tmp/sdr_ab.py 2048 > gpurun_out/sdr_simt.json 2> gpurun_out/sdr_simt.err; echo "simt rc=$?"; cat gpurun_out/sdr_simt.json
timeout 600 python tools/diag_parity_shards.py > gpurun_out/diag_shards.log 2>&1; echo "diag rc=$?"; cat gpurun_out/diag_shards.log | tail -10
