set -x
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_crafted.py -x -q -k "True" 2>&1 | tail -15
echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_crafted.py -x -q -k "True" 2>&1 | tail -25
