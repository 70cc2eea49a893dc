set -x
python -m pytest tests/test_gpu_pack.py -m gpu -x -q -s 2>&1 | tail -25
