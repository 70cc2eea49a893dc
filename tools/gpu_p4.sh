#!/bin/bash
# A/B of the split-parse grid (2 vs 3 CTAs per SM) on the batch one rank of an 8-way strong-scaled C2 sees
for e in 8192 4096 16384; do for c in 3 2; do
  echo "entries $e parse4 ctas $c"; ZPB_PARSE4_CTAS=$c python bench.py --entries $e --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(round(d['value'],1),'GB/s  step',round(d['ms_per_step'],3),'ms  all_kernels',round(r['all_kernels_ms'],3), r['stages_ms']['parse_ms'], r['stages_ms']['exec_ms'])"
done; done
