#!/bin/bash
# developer experiment: the round-1 library next to the current one, same box, same inputs
timeout 900 python -m pytest tests/test_gpu_unpack.py tests/test_gpu_crafted.py tests/test_gpu_blocks.py -x -q 2>&1 | tail -2
for lib in build_abl/v1.so zpack_b200/libzpack_b200.so; do
  echo "== $lib"
  ZPB_LIB=$PWD/$lib python tools/class_bench.py --entries 14208 --groups 8 --classes 1,2,3,-1 --reps 3 --overlap 0 2>&1 | cut -c1-330
done
