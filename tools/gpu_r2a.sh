#!/bin/bash
# round 2, first look at the rewritten parse / execute kernels: parity, per-class times, instruction counts
set -x
nproc; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_unpack.py tests/test_gpu_crafted.py tests/test_gpu_blocks.py -x -q 2>&1 | tail -5
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_crafted.py -x -q -k "True" 2>&1 | tail -8
python tools/class_bench.py --entries 14208 --groups 8 --classes 0,1,2,3,-1 --reps 3 --exec-ctas 3,4 --overlap 0,1 > gpurun_out/class_r2a.jsonl 2> gpurun_out/class_r2a.err
cat gpurun_out/class_r2a.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_fast_exec -s 2 -c 1 -o gpurun_out/exec_mixed_r2a \
    python bench.py --entries 28416 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_exec.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2a.csv \
    python bench.py --entries 16384 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2_r2a.json 2> gpurun_out/bench_c2_r2a.err; cat gpurun_out/bench_c2_r2a.json
