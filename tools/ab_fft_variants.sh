#!/bin/bash
# A/B of CTA size / min-blocks (register cap) variants of the two FFT kernels; run on the GPU box.
for v in "8 2" "4 4" "4 5" "8 3"; do
  set -- $v
  python -m fast_speech_enhancement_metrics_b200.build -DFSEM_FFT_WARPS=$1 -DFSEM_FFT_MINBLOCKS=$2 > /dev/null 2>&1
  python bench.py --batch 4096 --steps 3 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['kernels']; print('$v', 'step %.2f ms' % d['ms_per_step'], 'spectrum %.3f' % k['pesq_spectrum_kernel']['ms_per_step'], 'tob %.3f' % k['stoi_tob_kernel']['ms_per_step'])"
done
