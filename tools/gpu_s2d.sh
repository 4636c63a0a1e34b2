#!/bin/bash
# SPR parity tests (quick) + ncu --set full of the event-scan SPR kernels with source-level stall attribution
OUT=gpurun_out/${1:-s2d}; mkdir -p $OUT
K=${2:-spr_g2_scan_kernel|spr_g2_emit_kernel|spr_g2_prefix_kernel}
timeout 900 python -m pytest tests/test_gpu_spr.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "exit $?" >> $OUT/pytest.log; tail -5 $OUT/pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 12 -c 4 \
  -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-partitioned --no-mcmc --evals-per-step 2 --spr-batches-per-step 1 > $OUT/ncu.log 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/ncu_raw.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page source --csv > $OUT/ncu_source.csv 2>/dev/null
python tools/summarize_ncu.py $OUT/ncu_raw.csv > $OUT/ncu_summary.txt 2>&1
grep -E "Kernel Name|gpu__time_duration|dram__bytes_(read|write)|registers|issue_active|warps_active|inst_executed.sum|long_scoreboard|stalled_wait|lg_throttle|short_score" $OUT/ncu_summary.txt
