#!/bin/bash
# pack path check: parity tests, per-class throughput + ratio, one ncu capture of the block compressor
set -x
timeout 900 python -m pytest tests/test_gpu_pack.py tests/test_gpu_zstd.py -x -q 2>&1 | tail -5
python tools/pack_bench.py --entries 8192 --classes=-1,1,2,3,0 --reps 3 > gpurun_out/pack_$1.jsonl 2> gpurun_out/pack_$1.err
cat gpurun_out/pack_$1.jsonl; tail -3 gpurun_out/pack_$1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_pack_blocks -s 1 -c 1 -o gpurun_out/pack_blocks_$1 \
    python tools/pack_bench.py --entries 4096 --classes=-1 --reps 1 > gpurun_out/ncu_pack.log 2>&1
tail -2 gpurun_out/ncu_pack.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/pack_launches_$1.csv \
    python tools/pack_bench.py --entries 4096 --classes=-1 --reps 1 > /dev/null 2>&1
grep -c lz4_pack gpurun_out/pack_launches_$1.csv
