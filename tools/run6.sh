set -x
ncu --set full --clock-control none --import-source on -k regex:'lz4_fast_exec' -s 2 -c 1 -o gpurun_out/prof_rec2 python tools/class_bench.py --classes 3 --entries 16384 --groups 32 --reps 1 > gpurun_out/ncu_rec.log 2>&1
