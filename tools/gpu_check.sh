#!/bin/bash
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
timeout 300 python tools/e2e_breakdown.py 4 4 2>&1 | tail -3 | tee $OUT/e2e_breakdown.txt
timeout 300 python tools/e2e_breakdown.py 16 4 2>&1 | tail -3 | tee -a $OUT/e2e_breakdown.txt
timeout 600 python bench.py --no-cpu-baseline --spr-studies 64 > $OUT/bench64.json 2> $OUT/bench64.err; echo "bench exit $?"; tail -3 $OUT/bench64.err
python - <<PY
import json
d=json.load(open("$OUT/bench64.json"))
print("value",round(d["value"]), "logg_ms",round(d["ms_per_step"],4), "enqueue", round(d["host_enqueue_ms_per_step"],4), "gen_ms", round(d["loglik_general_schedule"]["launch_ms"],4), "spr_ms",round(d["spr_ms_per_batch"],4),"spr_c/s %.3g"%d["spr_candidates_per_s"], "regions", d["spr_regions_per_batch"], "e2e",round(d["e2e"]["value"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --spr-studies 64 > $OUT/bench_under_ncu.log 2>&1
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; grep -E "spr_|folded|kernel  " $OUT/launches_summary.txt | head -20
