ncu --set full --clock-control none --import-source on -k regex:lz4_fast_parse -s 2 -c 1 -o gpurun_out/s8h_parse python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/s8h_ncu.log 2>&1
