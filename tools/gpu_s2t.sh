#!/bin/bash
# full round + config-2 launch list + e2e thread sweep
tools/gpu_round.sh ${1:-s2t} stbln
OUT=gpurun_out/${1:-s2t}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_cfg2.csv \
    python bench.py --config 2 --chains 256 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-partitioned --no-mcmc --evals-per-step 2 --spr-studies 0 --e2e-chains 4 > $OUT/bench_cfg2_under_ncu.log 2>&1
python tools/summarize_launches.py $OUT/launches_cfg2.csv > $OUT/launches_cfg2_summary.txt 2>&1; grep -E "emat_|kernel  " $OUT/launches_cfg2_summary.txt
for t in 1 3 4; do timeout 300 python bench.py --e2e-threads $t --no-cpu-baseline --no-secondary --no-partitioned --no-mcmc --spr-studies 0 --steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e threads', $t, d['e2e']['value'])"; done
