set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc; lscpu | grep 'Model name'
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/class_bench.py --entries 4096 > gpurun_out/class_r1a.jsonl 2>gpurun_out/class_r1a.err
python bench.py --entries 16384 --steps 5 --warmup 3 > gpurun_out/bench_r1a.json 2>gpurun_out/bench_r1a.err
tail -3 gpurun_out/bench_r1a.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --entries 4096 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:unpack_kernel -s 2 -c 1 -o gpurun_out/prof_r1a python bench.py --entries 4096 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out
