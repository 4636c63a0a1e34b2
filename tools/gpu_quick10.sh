#!/bin/bash
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
for c in 0 1; do DPHY_FOLDED_CS=$c timeout 300 python tools/logg_occ.py 16 4 2>&1 | tail -2 | head -1 | sed "s/^/cs=$c /"; done | tee $OUT/logg_cs.txt
