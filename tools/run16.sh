set -x
nproc; lscpu | grep 'Model name'
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/s8_bench_c2.json 2>gpurun_out/s8_bench_c2.err; tail -3 gpurun_out/s8_bench_c2.err; cat gpurun_out/s8_bench_c2.json
python tools/class_bench.py --entries 8192 --groups 8 --classes 1,3,2 > gpurun_out/s8_class.jsonl 2>gpurun_out/s8_class.err; cat gpurun_out/s8_class.jsonl
ncu --set full --clock-control none --import-source on -k regex:lz4_fast_exec -s 2 -c 1 -o gpurun_out/s8_exec_text python tools/class_bench.py --entries 4096 --groups 8 --classes 1 --reps 1 > gpurun_out/s8_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lz4_fast_exec -s 2 -c 1 -o gpurun_out/s8_exec_rec python tools/class_bench.py --entries 4096 --groups 8 --classes 3 --reps 1 > gpurun_out/s8_ncu2.log 2>&1
ls -la gpurun_out
