set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/class_bench.py --entries 14208 --groups 8 --classes 1,3,2,-1 --reps 3 > gpurun_out/s8g_class.jsonl 2>gpurun_out/s8g_class.err; cut -c1-330 gpurun_out/s8g_class.jsonl; tail -3 gpurun_out/s8g_class.err
