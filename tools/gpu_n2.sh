#!/bin/bash
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_host_lib.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_c2_n2_r2.json 2> gpurun_out/bench_c2_n2_r2.err; tail -2 gpurun_out/bench_c2_n2_r2.err; cut -c1-1400 gpurun_out/bench_c2_n2_r2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c5 --entries 16384 --steps 5 --warmup 3 > gpurun_out/bench_c5_n2_r2.json 2> gpurun_out/bench_c5_n2_r2.err; tail -2 gpurun_out/bench_c5_n2_r2.err; cut -c1-600 gpurun_out/bench_c5_n2_r2.json
