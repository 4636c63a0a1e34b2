#!/bin/bash
# compute-sanitizer memcheck over the SPR / delta parity tests (small trees; the sanitizer slows kernels 10-50x)
OUT=gpurun_out/${1:-san}; mkdir -p $OUT
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_spr.py tests/test_gpu_delta.py -m gpu -x -q -k "${2:-grouped or reference_region or every_tree or detached or taken_as_given or enumerate or delta or thousand}" > $OUT/memcheck.log 2>&1; echo "memcheck exit $?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|at 0x|by thread" $OUT/memcheck.log | head -40
