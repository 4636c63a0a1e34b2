#!/bin/bash
OUT=gpurun_out/${1:-s3d}; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 600 --csv --log-file $OUT/launches_cfg2.csv \
    python bench.py --config 2 --chains 256 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-partitioned --no-mcmc --evals-per-step 2 --spr-studies 0 --e2e-chains 4 > $OUT/bench_cfg2_under_ncu.log 2>&1
python tools/summarize_launches.py $OUT/launches_cfg2.csv > $OUT/launches_cfg2_summary.txt 2>&1; grep -E "emat_|kernel  |flatten|delta" $OUT/launches_cfg2_summary.txt | head -30
