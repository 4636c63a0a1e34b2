#!/bin/bash
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
DPHY_FOLDED_PAIR=0 timeout 300 python tools/logg_occ.py 16 4 2>&1 | tail -2 | tee $OUT/logg_pair.txt
DPHY_FOLDED_PAIR=1 timeout 300 python tools/logg_occ.py 16 4 2>&1 | tail -2 | tee -a $OUT/logg_pair.txt
DPHY_FOLDED_PAIR=1 timeout 300 python tools/logg_occ.py 16 5 2>&1 | tail -2 | tee -a $OUT/logg_pair.txt
