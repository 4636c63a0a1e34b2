#!/bin/bash
OUT=gpurun_out/${1:-s3f}; mkdir -p $OUT
timeout 300 python tools/tally_timing.py 16 4 2>&1 | tail -6
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $OUT/launches_tally.csv python tools/tally_timing.py 16 4 > $OUT/tally_under_ncu.log 2>&1
python tools/summarize_launches.py $OUT/launches_tally.csv 2>&1 | grep -E "tally|kernel  |gather" | head -20
