bash tools/gpu_iter.sh ${1:-combo}
bash tools/gpu_sdr.sh 2>&1 | grep -v "^shard" | grep "passed\|failed\|simt"
