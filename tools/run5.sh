set -x
ncu --set full --clock-control none --import-source on -k regex:'lz4_fast_(parse|exec)' -s 2 -c 2 -o gpurun_out/prof_r1d python tools/class_bench.py --classes 1 --entries 16384 --groups 32 --reps 1 > gpurun_out/ncu_r1d.log 2>&1
tail -3 gpurun_out/ncu_r1d.log
