#!/bin/bash
# ncu launch list + full capture of selected kernels.  usage: tools/gpu_prof.sh <tag> <kernel-regex> [skip] [count]
TAG=${1:-p}; RX=${2:-emat_log_G_kernel}; SKIP=${3:-4}; CNT=${4:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT \
  -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
