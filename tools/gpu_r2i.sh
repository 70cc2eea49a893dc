#!/bin/bash
nproc; nvidia-smi -L | head -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default_r2i.json 2> gpurun_out/bench_default_r2i.err; tail -3 gpurun_out/bench_default_r2i.err; cut -c1-3000 gpurun_out/bench_default_r2i.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2i.json 2> gpurun_out/bench_ref_r2i.err; tail -3 gpurun_out/bench_ref_r2i.err; cut -c1-900 gpurun_out/bench_ref_r2i.json
