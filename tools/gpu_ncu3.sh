#!/bin/bash
export ZPB_OVERLAP=0 ZPB_PARSE_SPLIT=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_fast_parse4 -s 1 -c 1 -o gpurun_out/parse4_small \
    python tools/class_bench.py --entries 2048 --groups 8 --classes 1 --reps 1 --overlap 0 > gpurun_out/ncu_p4.log 2>&1
