#!/bin/bash
# bounded-study frontier kernel: parity (SPR + MCMC drop-in with every study verified) and timing
OUT=gpurun_out/${1:-s2y}; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_spr.py tests/test_gpu_mcmc.py tests/test_gpu_full_size.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "exit $?" >> $OUT/pytest.log; tail -8 $OUT/pytest.log
for lim in 1 2; do timeout 300 python tools/spr_bounded_timing.py $lim 2>&1 | tail -4; done
DPHY_SPR_FRONTIER=0 timeout 300 python tools/spr_bounded_timing.py 1 2>&1 | grep limit | head -1
