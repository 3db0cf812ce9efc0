#!/bin/bash
# A/B of alternative builds of the library: bash tools/gpu_ab_libs.sh lib1.so lib2.so ...  (paths relative to the repo)
mkdir -p gpurun_out
for lib in "$@"; do
  FSEM_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-e2e --no-cpu --no-parity --steps 5 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/ab.json"))
    print("$lib", "ms %.3f" % j["ms_per_step"], {k: round(v["ms_per_step"], 3) for k, v in j["kernels"].items() if v["ms_per_step"] > 0.05})
except Exception as e:
    print("$lib failed", e, open("gpurun_out/ab.err").read()[-400:])
PY
done
