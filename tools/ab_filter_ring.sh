#!/bin/bash
# A/B of L2 prefetch policies for the IIR pass's row-strided tile reads; run on the GPU box.
for v in "-DFSEM_FILTER_L2_256" "-DFSEM_FILTER_PF=128" "-DFSEM_FILTER_PF=256" "-DFSEM_FILTER_PF=128 -DFSEM_FILTER_L2_256" "-DFSEM_FILTER_PF=64"; do
  python -m fast_speech_enhancement_metrics_b200.build $v > /dev/null 2>&1
  for b in 8192 1024; do
  python bench.py --batch $b --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('[$v] batch $b', 'step %.2f ms' % d['ms_per_step'], 'filter %.3f' % k['pesq_filter_kernel']['ms_per_step'])"
  done
done
python -m fast_speech_enhancement_metrics_b200.build > /dev/null 2>&1
