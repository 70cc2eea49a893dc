#!/bin/bash
python tools/class_bench.py --entries 14208 --groups 8 --classes 1,3 --reps 3 --overlap 0 2>&1 | cut -c60-500
export ZPB_OVERLAP=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_fast_parse4 -s 1 -c 1 -o gpurun_out/parse4_$1 \
    python bench.py --entries 28416 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_parse.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_fast_exec -s 1 -c 1 -o gpurun_out/exec_mixed_$1 \
    python bench.py --entries 28416 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_exec.log 2>&1
