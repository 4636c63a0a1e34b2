#!/bin/bash
OUT=gpurun_out/${1:-s3e}; mkdir -p $OUT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_spr.py -m gpu -x -q -k "synthetic_studies" > $OUT/memcheck.log 2>&1; echo "memcheck exit $?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|at 0x" $OUT/memcheck.log | head -20
tools/gpu_s2c.sh $1
