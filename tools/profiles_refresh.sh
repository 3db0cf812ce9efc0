#!/bin/bash
# Copy the outputs of `gpurun -- bash tools/gpu_evidence.sh TAG` from gpurun_out/ into profiles/ under the round's names.
#   bash tools/profiles_refresh.sh r02b r02
TAG=${1:?evidence tag}; RND=${2:-r02}
G=gpurun_out; P=profiles
cp $G/${TAG}final_bench.json $P/${RND}_bench_1gpu.json
cp $G/${TAG}final_bench_ref.json $P/${RND}_bench_reference_arm.json
python tools/ncu_launches_summary.py $G/${TAG}final_launches.csv $P/${RND}_ncu_launches > /dev/null
python tools/ncu_raw_summary.py $G/${TAG}_full_all_raw.csv --md $P/${RND}_ncu_full.md --json $P/${RND}_ncu_pipes.json > /dev/null
cp $G/${TAG}_config_sweep.json $P/${RND}_config_sweep.json
python tools/sass_opcodes.py --md $P/${RND}_sass_opcodes.md > /dev/null || echo "sass_opcodes failed"
tail -3 $G/${TAG}final_pytest.log
ls -la $P/${RND}_*
