#!/bin/bash
# strong scaling of the contract bench at N GPUs (one archive of the config's size cut over the ranks), its reference
# arm, the product's own multi-GPU call, and the block-sharded entry.  Usage: gpurun --gpus N -- 'bash tools/gpu_scale.sh N'
N=$1
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -x -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_c2_n$N.json 2> gpurun_out/r2_bench_c2_n$N.err; tail -2 gpurun_out/r2_bench_c2_n$N.err; cut -c1-1300 gpurun_out/r2_bench_c2_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --scaling weak --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_bench_c2_weak_n$N.json 2> gpurun_out/r2_bench_c2_weak_n$N.err; cut -c1-300 gpurun_out/r2_bench_c2_weak_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c5 --steps 5 --warmup 3 > gpurun_out/r2_bench_c5_n$N.json 2> gpurun_out/r2_bench_c5_n$N.err; tail -2 gpurun_out/r2_bench_c5_n$N.err; cut -c1-500 gpurun_out/r2_bench_c5_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err; cut -c1-300 gpurun_out/r2_bench_c3_n$N.json
