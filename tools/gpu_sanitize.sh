#!/bin/bash
# compute-sanitizer memcheck + racecheck over the pack path (LZ4 and zstd writers) and a small unpack, through the tests
export ZPB_HOST_WORKERS=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pack.py -x -q -k "oracle or zstd_ratio or none_method" 2>&1 | tail -8
echo "memcheck rc=$?"
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report hazard --error-exitcode 9 python -m pytest tests/test_gpu_pack.py -x -q -k "none_method" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|Hazard|hazard" | head -12
