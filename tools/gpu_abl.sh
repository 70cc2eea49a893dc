#!/bin/bash
# developer experiment: the execute kernel with one phase switched off (wrong output on purpose) — what each phase costs
for v in DEP LIT MATCH; do
  echo "== without $v"
  ZPB_NOCHECK=1 ZPB_LIB=$PWD/build_abl/abl_$v.so python tools/class_bench.py --entries 14208 --groups 8 --classes 1,3 --reps 3 --exec-ctas 3 --overlap 0 2>&1 | cut -c1-330
done
