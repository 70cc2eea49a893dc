set -x
python tools/pack_bench.py --entries 8192 --classes 0,1,2,3,-1 > gpurun_out/pack_r1a.jsonl 2> gpurun_out/pack_r1a.err
cat gpurun_out/pack_r1a.jsonl; tail -3 gpurun_out/pack_r1a.err
