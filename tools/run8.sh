set -x
# launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --entries 16384 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/b_ncu_r1.log 2>&1
# full capture of the dominant kernel (one launch) on the same command
ncu --set full --clock-control none --import-source on -k regex:lz4_fast_exec -s 3 -c 1 -o gpurun_out/prof_r1_exec python bench.py --entries 16384 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/b_ncu2_r1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lz4_fast_parse -s 3 -c 1 -o gpurun_out/prof_r1_parse python bench.py --entries 16384 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/b_ncu3_r1.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1c.json 2>gpurun_out/bench_r1c.err
cat gpurun_out/bench_r1c.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1c_ref.json 2>gpurun_out/bench_r1c_ref.err
cat gpurun_out/bench_r1c_ref.json
