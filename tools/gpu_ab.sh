#!/bin/bash
for lib in build_abl/v1.so build_abl/V1.so build_abl/V2.so zpack_b200/libzpack_b200.so; do
  echo "== $lib"
  ZPB_NOCHECK=1 ZPB_LIB=$PWD/$lib python tools/class_bench.py --entries 14208 --groups 8 --classes 1,-1 --reps 3 --overlap 0 2>&1 | cut -c60-330
done
