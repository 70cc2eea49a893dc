#!/bin/bash
# round-2 evidence for the pack path: all GPU tests, C3 bench + its reference arm, per-class throughput and ratio,
# launch list and --set full captures of the two pack kernels
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; cat gpurun_out/r2_bench_c3.json; tail -2 gpurun_out/r2_bench_c3.err
python bench.py --workload c3 --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_c3_reference.json 2> gpurun_out/r2_bench_c3_reference.err; cat gpurun_out/r2_bench_c3_reference.json
python tools/pack_bench.py --entries 16384 --classes=-1,1,2,3,0 --reps 3 > gpurun_out/r2_pack_class_bench.jsonl 2> gpurun_out/pack.err; cat gpurun_out/r2_pack_class_bench.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_pack_c3_8192.csv \
    python bench.py --workload c3 --entries 8192 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch_pack.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_pack_blocks -s 1 -c 1 -o gpurun_out/r2_pack_blocks_mixed \
    python tools/pack_bench.py --entries 8192 --classes=-1 --reps 1 > gpurun_out/ncu_pack.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_pack_blocks -s 1 -c 1 -o gpurun_out/r2_pack_blocks_text \
    python tools/pack_bench.py --entries 8192 --classes=1 --reps 1 > gpurun_out/ncu_pack2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_pack_kernel -s 1 -c 1 -o gpurun_out/r2_pack_frames_mixed \
    python tools/pack_bench.py --entries 8192 --classes=-1 --reps 1 > gpurun_out/ncu_pack3.log 2>&1
ls -la gpurun_out | tail -12
