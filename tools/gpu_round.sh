#!/bin/bash
# One iteration call: GPU tests on the in-tree library, A/B bench of alternative builds, small-batch sweep (graph path).
#   gpurun --timeout 1200 -- bash tools/gpu_round.sh tag lib_base.so lib_x.so ...
TAG=${1:-it}; shift
mkdir -p gpurun_out
timeout 700 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
bash tools/gpu_ab_libs.sh "$@" fast_speech_enhancement_metrics_b200/libfsem_b200.so 2>&1 | tee gpurun_out/${TAG}_ab.log
timeout 300 python tools/config_sweep.py --max-batch 1024 --out gpurun_out/${TAG}_sweep_small.json > gpurun_out/${TAG}_sweep_small.log 2>&1
echo "sweep rc=$?"; cat gpurun_out/${TAG}_sweep_small.log | tail -20
