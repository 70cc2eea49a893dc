set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/class_bench.py --method zstd --entries 4096 --groups 32 --reps 3 2>&1 | tee gpurun_out/zstd_class_r1.jsonl | tail -8
