#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_unpack.py tests/test_gpu_crafted.py tests/test_gpu_blocks.py -x -q 2>&1 | tail -2
for sp in 0 1; do
echo "== ZPB_PARSE_SPLIT=$sp"
for n in 2048 8192 16384; do
ZPB_PARSE_SPLIT=$sp python tools/class_bench.py --entries $n --groups 8 --classes 1,3,-1 --reps 5 --overlap 1 2>&1 | cut -c60-420
done
done
