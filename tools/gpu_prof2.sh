#!/bin/bash
# tests, then ncu launch list + full capture of selected kernels using tools/logg_timing-like driver (mask 0 only)
TAG=${1:-p}; RX=${2:-emat_log_G}; SKIP=${3:-4}; CNT=${4:-3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -3 $OUT/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --spr-studies 0 > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT \
  -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --spr-studies 0 > $OUT/ncu_full.log 2>&1
ls -la $OUT
