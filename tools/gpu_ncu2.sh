#!/bin/bash
export ZPB_OVERLAP=0
for v in v1 cur; do
  lib=zpack_b200/libzpack_b200.so; [ $v = v1 ] && lib=build_abl/v1.so
  ZPB_LIB=$PWD/$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_fast_exec -s 1 -c 1 -o gpurun_out/exec_ab_$v \
    python tools/class_bench.py --entries 7104 --groups 8 --classes 1 --reps 1 --overlap 0 > gpurun_out/ncu_ab_$v.log 2>&1
done
