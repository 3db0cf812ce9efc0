#!/bin/bash
# A/B of IIR-pass variants (compile-time switches); run on the GPU box.  Leaves the default build in place.
for v in "" "-DFSEM_FILTER_DRAIN32" "-DFSEM_FILTER_RING"; do
  python -m fast_speech_enhancement_metrics_b200.build $v > /dev/null 2>&1
  for b in 8192 8192 1024; do
  python bench.py --batch $b --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('[$v] batch $b', 'step %.2f ms' % d['ms_per_step'], 'filter %.3f' % k['pesq_filter_kernel']['ms_per_step'], 'spectrum %.3f' % k['pesq_spectrum_kernel']['ms_per_step'])"
  done
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
done
python -m fast_speech_enhancement_metrics_b200.build > /dev/null 2>&1
