set -x
python -m pytest tests/test_host_lib.py -m gpu -x -q 2>&1 | tail -30
