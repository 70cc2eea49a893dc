#!/bin/bash
# round-2 evidence in one call (one GPU): all GPU tests, smoke, the contract bench + reference arm, C4/C5 lines,
# C1 through both CLIs, launch list of the bench command, --set full captures of the kernels the verdict names
set -x
nproc; lscpu | grep 'Model name'
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_c2_reference.json 2> gpurun_out/r2_bench_c2_reference.err; cat gpurun_out/r2_bench_c2_reference.json
SECONDS=0; python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; echo "default bench.py wall time: ${SECONDS} s"; cut -c1-1500 gpurun_out/r2_bench_c2.json; tail -2 gpurun_out/r2_bench_c2.err
python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; cut -c1-600 gpurun_out/r2_bench_c4.json
python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; cut -c1-400 gpurun_out/r2_bench_c3.json
python bench.py --workload c3z --steps 5 --warmup 3 > gpurun_out/r2_bench_c3z.json 2> gpurun_out/r2_bench_c3z.err; cut -c1-400 gpurun_out/r2_bench_c3z.json
python bench.py --workload c3z --impl reference --entries 8192 --steps 2 --warmup 1 > gpurun_out/r2_bench_c3z_reference.json 2>/dev/null
python tools/pack_bench.py --entries 8192 --classes=-1,1,2,3,0 --reps 3 > gpurun_out/r2_pack_class_bench.jsonl 2>/dev/null; python tools/pack_bench.py --entries 8192 --classes=-1,1,2,3,0 --reps 3 --method 1 > gpurun_out/r2_zstd_pack_class_bench.jsonl 2>/dev/null; cat gpurun_out/r2_pack_class_bench.jsonl gpurun_out/r2_zstd_pack_class_bench.jsonl
python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err; cut -c1-600 gpurun_out/r2_bench_c5.json
python tools/c1_cli.py > gpurun_out/r2_c1_cli.jsonl 2> gpurun_out/r2_c1_cli.err; cat gpurun_out/r2_c1_cli.jsonl
python tools/class_bench.py --entries 14208 --groups 8 --classes 0,1,2,3,-1 --reps 3 > gpurun_out/r2_class_bench.jsonl 2> gpurun_out/class.err; cut -c1-300 gpurun_out/r2_class_bench.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_bench_16384.csv \
    python bench.py --entries 16384 --steps 2 --warmup 1 --no-e2e --no-cpu --no-configs > gpurun_out/ncu_launch.log 2>&1
ZPB_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_fast_exec -s 1 -c 1 -o gpurun_out/r2_exec_mixed28416 \
    python bench.py --entries 28416 --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs > gpurun_out/ncu_exec.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lz4_fast_parse4 -s 1 -c 1 -o gpurun_out/r2_parse4_8192 \
    python bench.py --entries 8192 --steps 2 --warmup 1 --no-cpu --no-e2e --no-configs > gpurun_out/ncu_parse.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zstd_unpack -s 1 -c 1 -o gpurun_out/r2_zstd_c4_8192 \
    python bench.py --workload c4 --entries 8192 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_zstd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xxh3_chain -s 1 -c 1 -o gpurun_out/r2_chain_c5 \
    python bench.py --workload c5 --entries 8192 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_chain.log 2>&1
python tools/archive_bench.py > gpurun_out/r2_archive_bench.jsonl 2> gpurun_out/r2_archive_bench.err; cat gpurun_out/r2_archive_bench.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_archive.csv python tools/archive_bench.py 16384 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:arc_copy -s 3 -c 1 -o gpurun_out/r2_arc_copy python tools/archive_bench.py 16384 > gpurun_out/ncu_arc.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_archive.py -x -q -k "not file_to_device" 2>&1 | tail -4
ls -la gpurun_out | tail -25
