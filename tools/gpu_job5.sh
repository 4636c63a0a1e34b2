#!/bin/bash
OUT=gpurun_out/${1:-j5}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_spr.py tests/test_gpu_full_size.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "exit $?" >> $OUT/pytest.log; tail -8 $OUT/pytest.log
tools/gpu_round.sh ${1:-j5}n l > $OUT/round.log 2>&1; grep -E "spr_|Kernel" gpurun_out/${1:-j5}n/launches_summary.txt
