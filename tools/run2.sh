set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/class_bench.py --entries 4096 --groups 32 > gpurun_out/class_r1b.jsonl 2>gpurun_out/class_r1b.err
cat gpurun_out/class_r1b.jsonl; tail -5 gpurun_out/class_r1b.err
