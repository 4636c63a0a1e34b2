#!/bin/bash
# One GPU call: parity tests, bench line, ncu launch list, one full ncu capture of the hot kernels.
# usage: tools/gpu_round.sh <tag>
TAG=${1:-rX}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cat $OUT/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref exit $?"
cat $OUT/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'emat_log_G_stream_kernel|emat_log_G_tile_kernel|spr_scan_kernel|spr_emit_kernel' -s 8 -c 4 \
  -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --chains 16 > $OUT/ncu_full.log 2>&1
ls -la $OUT
