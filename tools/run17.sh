set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/class_bench.py --entries 14208 --groups 8 --classes 2 --reps 3 > gpurun_out/s8i_class.jsonl 2>gpurun_out/s8i_class.err; cut -c1-330 gpurun_out/s8i_class.jsonl; tail -3 gpurun_out/s8i_class.err
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['stages_ms'])"
