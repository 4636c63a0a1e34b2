#!/bin/bash
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python tools/logg_timing.py 16 4 2>&1 | tail -9 | tee $OUT/logg_masks.txt
timeout 300 python tools/logg_timing.py 16 2 2>&1 | tail -9 | tee $OUT/logg_masks_cfg2.txt
