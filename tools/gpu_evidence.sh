#!/bin/bash
# Round evidence in one call: full check (tests, bench, reference arm, ncu launch list) + ncu --set full of the step's kernels
# + config sweep.    gpurun --timeout 2400 -- bash tools/gpu_evidence.sh r02
TAG=${1:-r02}
bash tools/gpu_check.sh ${TAG}final
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pesq_spectrum|pesq_filter_tiled|pesq_bark|stoi_resample85|stoi_tob|stoi_segment" \
    -s 18 -c 6 -o gpurun_out/${TAG}_full_all -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity --no-graph > gpurun_out/${TAG}_ncufull_all.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/${TAG}_full_all.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_all_raw.csv 2>/dev/null
timeout 900 python tools/config_sweep.py --out gpurun_out/${TAG}_config_sweep.json > gpurun_out/${TAG}_config_sweep.log 2>&1
echo "sweep rc=$?"; tail -25 gpurun_out/${TAG}_config_sweep.log
