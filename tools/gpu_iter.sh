#!/bin/bash
# Quick iteration call: GPU parity tests + device-timed bench (no e2e / cpu legs); NCU="regex" adds one --set full capture.
#   gpurun --timeout 900 -- bash tools/gpu_iter.sh tag [ncu-kernel-regex]
TAG=${1:-it}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --no-e2e --no-cpu --steps 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
python - <<PY
import json
try:
    j = json.load(open("gpurun_out/${TAG}_bench.json"))
    print("value %.0f ms %.3f" % (j["value"], j["ms_per_step"]))
    print("kernels", {k: round(v["ms_per_step"], 3) for k, v in j["kernels"].items() if v["ms_per_step"] > 0})
    p = j["parity"]; print("parity ok", p["ok"], p["pesq_max_abs"], p["stoi_max_abs"], p["estoi_max_abs"], p.get("vs_reference"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${TAG}_bench.err").read()[-1500:])
PY
if [ -n "$2" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s 3 -c 1 -o gpurun_out/${TAG}_full -f \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity --no-graph > gpurun_out/${TAG}_ncufull.log 2>&1
  echo "ncu full rc=$?"
  ncu -i gpurun_out/${TAG}_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
fi
