#!/bin/bash
# one --set full capture of the execute kernel (and optionally the parse kernel) on the mixed 28416-entry batch;
# ZPB_OVERLAP=0 so that the execute kernel is ONE launch per step (under ncu kernels are serialised anyway)
set -x
export ZPB_OVERLAP=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_fast_exec -s 1 -c 1 -o gpurun_out/exec_mixed_$1 \
    python bench.py --entries 28416 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_exec.log 2>&1
if [ -n "$2" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_fast_parse -s 1 -c 1 -o gpurun_out/parse_mixed_$1 \
    python bench.py --entries 28416 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_parse.log 2>&1
fi
tail -2 gpurun_out/ncu_exec.log
