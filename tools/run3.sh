set -x
ncu --set full --clock-control none --import-source on -k regex:'lz4_fast_(parse|exec)' -s 2 -c 2 -o gpurun_out/prof_r1b python tools/class_bench.py --classes 1 --entries 16384 --groups 32 --reps 1 > gpurun_out/ncu_r1b.log 2>&1
tail -3 gpurun_out/ncu_r1b.log
python tools/class_bench.py --entries 16384 --groups 32 --classes 1,3,-1 > gpurun_out/class_r1c.jsonl 2>gpurun_out/class_r1c.err
cat gpurun_out/class_r1c.jsonl; tail -5 gpurun_out/class_r1c.err
