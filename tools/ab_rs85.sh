#!/bin/bash
# A/B of blocks per thread / tiles per CTA in the 8:5 resampler; run on the GPU box.
for v in "2 4" "2 8" "2 16"; do
  set -- $v
  python -m fast_speech_enhancement_metrics_b200.build -DFSEM_RS85_BLOCKS=$1 -DFSEM_RS85_TILES=$2 > /dev/null 2>&1
  python bench.py --batch 4096 --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('NB=$1 tiles=$2', 'step %.2f ms' % d['ms_per_step'], 'resample %.3f' % k['stoi_resample_kernel']['ms_per_step'])"
done
