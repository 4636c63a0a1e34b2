#!/bin/bash
# bench (new e2e legs) + ncu of the grouped SPR kernels with source-level stall attribution
OUT=gpurun_out/${1:-s2b}; mkdir -p $OUT
timeout 900 python bench.py --no-secondary --no-mcmc --no-partitioned > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
for k in ("value","e2e","e2e_edit","spr_ms_per_batch","model_change_cycle"): print(k, d.get(k))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spr_gscan_kernel|spr_gemit_kernel" -s 6 -c 2 \
  -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-partitioned --no-mcmc --evals-per-step 2 --spr-batches-per-step 1 > $OUT/ncu.log 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/ncu_raw.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page source --csv > $OUT/ncu_source.csv 2>/dev/null
python tools/summarize_ncu.py $OUT/ncu_raw.csv > $OUT/ncu_summary.txt 2>&1
grep -E "Kernel Name|gpu__time_duration|dram__bytes_(read|write)|registers|issue_active|warps_active|inst_executed.sum" $OUT/ncu_summary.txt
