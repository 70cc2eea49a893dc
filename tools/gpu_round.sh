#!/bin/bash
# What a round's GPU evidence is made of, in one `gpurun` call (one GPU):
#   gpurun --timeout 1800 -- 'bash tools/gpu_round.sh'
# parity tests, the contract bench + its reference arm, per-class numbers, the ncu launch list of the bench command
# and one `--set full` capture of each of the two LZ4 kernels.  Everything lands in gpurun_out/; the summaries that
# are meant to be judged are then copied / exported into profiles/ (tools/ncu_summary.py, tools/ncu_lines.py).
set -x
nproc; lscpu | grep 'Model name'
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_ref.json 2> gpurun_out/bench_c2_ref.err; cat gpurun_out/bench_c2_ref.json
python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json
python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; cat gpurun_out/bench_c5.json
python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
python tools/class_bench.py --entries 14208 --groups 8 --classes 0,1,2,3,-1 --reps 3 > gpurun_out/class.jsonl 2> gpurun_out/class.err
python tools/pack_bench.py > gpurun_out/pack.jsonl 2> gpurun_out/pack.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --entries 16384 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lz4_fast_exec -s 2 -c 1 -o gpurun_out/exec_mixed \
    python bench.py --entries 28416 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_exec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lz4_fast_parse -s 2 -c 1 -o gpurun_out/parse_c2 \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_parse.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
