#!/bin/bash
# quick look at a kernel change: LZ4 parity tests, per-class times, one ncu capture of the execute kernel
set -x
timeout 900 python -m pytest tests/test_gpu_unpack.py tests/test_gpu_crafted.py tests/test_gpu_blocks.py -x -q 2>&1 | tail -3
python tools/class_bench.py --entries 14208 --groups 8 --classes 1,2,3,-1 --reps 3 --exec-ctas ${CTAS:-3} --overlap 0 > gpurun_out/class_$1.jsonl 2> gpurun_out/class_$1.err
cut -c1-400 gpurun_out/class_$1.jsonl
export ZPB_OVERLAP=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_fast_exec -s 1 -c 1 -o gpurun_out/exec_mixed_$1 \
    python bench.py --entries 28416 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_exec.log 2>&1
