#!/bin/bash
# Quick GPU check: parity tests + bench lines at two SPR batch sizes + launch list.  usage: tools/gpu_quick2.sh <tag>
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench16.json 2> $OUT/bench16.err; echo "bench exit $?"; tail -3 $OUT/bench16.err
timeout 600 python bench.py --no-cpu-baseline --spr-studies 64 > $OUT/bench64.json 2> $OUT/bench64.err; echo "bench exit $?"; tail -3 $OUT/bench64.err
python - <<PY
import json
for n in ("16","64"):
    try:
        d=json.load(open("$OUT/bench%s.json"%n))
        print(n, "value",round(d["value"]), "logg_ms",round(d["ms_per_step"],4), "gen_ms", round(d["loglik_general_schedule"]["launch_ms"],4), "spr_ms",round(d["spr_ms_per_batch"],4),"spr_c/s %.3g"%d["spr_candidates_per_s"], "regions", d["spr_regions_per_batch"], "e2e",round(d["e2e"]["value"]))
    except Exception as e: print(n, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; grep -E "spr_|folded|kernel  " $OUT/launches_summary.txt | head -20
