set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/class_bench.py --entries 16384 --groups 32 > gpurun_out/class_r1d.jsonl 2>gpurun_out/class_r1d.err
cat gpurun_out/class_r1d.jsonl; tail -5 gpurun_out/class_r1d.err
