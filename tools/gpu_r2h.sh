#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for sp in 0 1; do
echo "== ZPB_PARSE_SPLIT=$sp"
ZPB_PARSE_SPLIT=$sp python tools/class_bench.py --entries 14208 --groups 8 --classes 1,-1 --reps 3 --overlap 0,1 2>&1 | cut -c60-330
ZPB_PARSE_SPLIT=$sp python tools/class_bench.py --entries 2048 --groups 8 --classes 1,-1 --reps 3 --overlap 0,1 2>&1 | cut -c60-330
done
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2_r2h.json 2> gpurun_out/bench_c2_r2h.err; cut -c1-120 gpurun_out/bench_c2_r2h.json; grep -o '"stages_ms": {[^}]*}' gpurun_out/bench_c2_r2h.json; grep -o '"all_kernels_ms": [0-9.]*' gpurun_out/bench_c2_r2h.json
