#!/bin/bash
# One gpurun call: GPU tests, bench (b200 + reference arm), ncu launch list of one step.  Logs under gpurun_out/.
#   gpurun --timeout 1500 -- bash tools/gpu_check.sh [tag]
TAG=${1:-chk}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt; free -g >> gpurun_out/${TAG}_gpu.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"pesq_|stoi_" -s 30 -c 10 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity --no-graph > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
python - <<PY
import json
try:
    j = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"] if j["e2e"] else None)
    print("kernels", {k: round(v["ms_per_step"], 3) for k, v in j["kernels"].items() if v["ms_per_step"] > 0})
    print("parity", j["parity"])
    print("cpu", j["cpu_baseline"])
except Exception as e:
    print("bench parse failed", e)
try:
    print("ref", open("gpurun_out/${TAG}_bench_ref.json").read()[:600])
except Exception as e:
    print(e)
PY
