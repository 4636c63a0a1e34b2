#!/bin/bash
# tests + bench (default) + bench (64 studies) + reference arm + launch list
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -8 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
timeout 600 python bench.py --no-cpu-baseline --spr-studies 64 > $OUT/bench64.json 2> $OUT/bench64.err; echo "bench exit $?"; tail -3 $OUT/bench64.err
python - <<PY
import json
for n in ("","64"):
    d=json.load(open("$OUT/bench%s.json"%n))
    print(n,"value",round(d["value"]), "logg_ms",round(d["ms_per_step"],4), "gen_ms", round(d["loglik_general_schedule"]["launch_ms"],4), "spr_ms",round(d["spr_ms_per_batch"],4),"spr_c/s %.3g"%d["spr_candidates_per_s"], "regions", d["spr_regions_per_batch"], "e2e",round(d["e2e"]["value"]))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_arm.json 2> $OUT/bench_ref.err; echo "ref exit $?"; cut -c1-400 $OUT/bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --spr-studies 64 > $OUT/bench_under_ncu.log 2>&1
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; grep -E "spr_|folded|flatten_ev|fold_br|kernel  " $OUT/launches_summary.txt | head -20
