python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/s8_bench_c5.json 2>gpurun_out/s8_bench_c5.err; tail -3 gpurun_out/s8_bench_c5.err; cat gpurun_out/s8_bench_c5.json
python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/s8_bench_c4.json 2>gpurun_out/s8_bench_c4.err; tail -3 gpurun_out/s8_bench_c4.err; cat gpurun_out/s8_bench_c4.json
python tools/pack_bench.py > gpurun_out/s8_pack.jsonl 2>gpurun_out/s8_pack.err; cat gpurun_out/s8_pack.jsonl; tail -2 gpurun_out/s8_pack.err
