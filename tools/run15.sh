set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2>gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2>gpurun_out/bench_c4.err; tail -3 gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json
python bench.py --workload c4 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c4_ref.json 2>gpurun_out/bench_c4_ref.err; cat gpurun_out/bench_c4_ref.json
