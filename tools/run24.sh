set -x
nproc; lscpu | grep 'Model name'
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/f_bench_c2.json 2>gpurun_out/f_bench_c2.err; tail -3 gpurun_out/f_bench_c2.err; cat gpurun_out/f_bench_c2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_c2_ref.json 2>gpurun_out/f_bench_c2_ref.err; cat gpurun_out/f_bench_c2_ref.json
python tools/class_bench.py --entries 14208 --groups 8 --classes 0,1,2,3,-1 --reps 3 > gpurun_out/f_class.jsonl 2>gpurun_out/f_class.err; cut -c1-330 gpurun_out/f_class.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/f_launches.csv python bench.py --entries 16384 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/f_ncu_launch.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
