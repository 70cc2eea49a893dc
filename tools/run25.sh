free -g | head -2; nproc
time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 > gpurun_out/s8_bench_c2_n8.json 2>gpurun_out/s8_bench_c2_n8.err; tail -3 gpurun_out/s8_bench_c2_n8.err; cat gpurun_out/s8_bench_c2_n8.json
time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --impl reference > gpurun_out/s8_bench_c2_n8_ref.json 2>gpurun_out/s8_bench_c2_n8_ref.err; tail -2 gpurun_out/s8_bench_c2_n8_ref.err; cat gpurun_out/s8_bench_c2_n8_ref.json
