#!/bin/bash
# SPR parity tests + a launch list of the bench's SPR batches.  usage: tools/gpu_s2c.sh <tag> [pytest args]
OUT=gpurun_out/${1:-s2c}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_spr.py tests/test_gpu_full_size.py tests/test_gpu_delta.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "exit $?" >> $OUT/pytest.log; tail -15 $OUT/pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-partitioned --no-mcmc --evals-per-step 2 --spr-batches-per-step 1 > $OUT/bench_under_ncu.log 2>&1
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; grep -E "spr_|kernel  " $OUT/launches_summary.txt
timeout 600 python bench.py --no-cpu-baseline --no-secondary --no-partitioned --no-mcmc > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
for k in ("value","e2e","e2e_edit","spr_ms_per_batch","spr_candidates_per_s","roofline_spr"): print(k, d.get(k))
PY
