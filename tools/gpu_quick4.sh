#!/bin/bash
# tests + folded occupancy sweep + bench (64 studies) + ncu full of the SPR emit / scan kernels and the folded log-G kernel
TAG=${1:-q}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -8 $OUT/pytest_gpu.log
for o in 4 5 6; do DPHY_FOLDED_OCC=$o timeout 300 python tools/logg_occ.py 16 4 2>&1 | tail -2 | head -1; done | tee $OUT/logg_occ.txt
timeout 600 python bench.py --no-cpu-baseline --spr-studies 64 > $OUT/bench64.json 2> $OUT/bench64.err; echo "bench exit $?"; tail -3 $OUT/bench64.err
python - <<PY
import json
d=json.load(open("$OUT/bench64.json"))
print("value",round(d["value"]), "logg_ms",round(d["ms_per_step"],4), "enqueue", round(d["host_enqueue_ms_per_step"],4), "gen_ms", round(d["loglik_general_schedule"]["launch_ms"],4), "spr_ms",round(d["spr_ms_per_batch"],4),"spr_c/s %.3g"%d["spr_candidates_per_s"], "regions", d["spr_regions_per_batch"], "e2e",round(d["e2e"]["value"]))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --spr-studies 64 > $OUT/bench_under_ncu.log 2>&1
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; grep -E "spr_|folded|flatten_ev|fold_br|kernel  " $OUT/launches_summary.txt | head -20
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spr_emit_kernel|spr_scan_kernel|emat_log_G_folded_kernel" -s 9 -c 3 \
    -o $OUT/prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline --spr-studies 16 > $OUT/ncu_full.log 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/ncu_raw.csv 2>/dev/null
python tools/summarize_ncu.py $OUT/ncu_raw.csv > $OUT/ncu_summary.txt 2>&1; grep -E "Kernel Name|duration|dram__bytes_(read|write)" $OUT/ncu_summary.txt
