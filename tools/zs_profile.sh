# developer tool: per-phase cycle shares of zstd_unpack_kernel (profiling build of the library)
set -x
ZPB_LIB=$PWD/zpack_b200/libzpack_b200_prof.so timeout 600 python tools/class_bench.py --method zstd --entries ${1:-8192} --groups 32 --reps 3 --classes ${2:-1,3,-1} 2>&1 | tee gpurun_out/zstd_profile.jsonl | tail -12
