set -x
timeout 600 python -m pytest tests/test_gpu_zstd.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/class_bench.py --method zstd --entries 8192 --groups 32 --reps 3 --classes 1,3,-1 2>&1 | tail -4
bash tools/zs_profile.sh 8192 1,3,-1
