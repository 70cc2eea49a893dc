#!/bin/bash
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/class_bench.py --entries 14208 --groups 8 --classes 1,2,3,-1 --reps 3 --overlap 0,1 2>&1 | cut -c1-330
python tools/class_bench.py --entries 2048 --groups 8 --classes 1,-1 --reps 3 --overlap 0,1 2>&1 | cut -c1-330
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2_r2g.json 2> gpurun_out/bench_c2_r2g.err; cut -c1-1500 gpurun_out/bench_c2_r2g.json
