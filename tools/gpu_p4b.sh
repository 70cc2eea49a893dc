#!/bin/bash
run() { python bench.py --entries 8192 --steps 10 --warmup 3 --no-e2e --no-cpu --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print(round(d['value'],1),'GB/s  step',round(d['ms_per_step'],3),'ms  all_kernels',round(r['all_kernels_ms'],3), round(r['stages_ms']['parse_ms'],3), round(r['stages_ms']['exec_ms'],3))"; }
echo "default"; run
echo "overlap 0"; ZPB_OVERLAP=0 run
echo "p4=1"; ZPB_PARSE4_CTAS=1 run
echo "p4=1 early=1"; ZPB_PARSE4_CTAS=1 ZPB_EARLY_CTAS=1 run
echo "p4=2 early=1"; ZPB_PARSE4_CTAS=2 ZPB_EARLY_CTAS=1 run
echo "nosplit"; ZPB_PARSE_SPLIT=0 run
python tools/class_bench.py --entries 8192 --groups 8 --classes 2,-1 --reps 3 --overlap 1 2>/dev/null | cut -c1-400
