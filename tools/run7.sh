set -x
nproc; free -g | head -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1b.json 2>gpurun_out/bench_r1b.err
cat gpurun_out/bench_r1b.json; tail -5 gpurun_out/bench_r1b.err
