#!/bin/bash
OUT=gpurun_out/${1:-s3k}; mkdir -p $OUT
true
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 700 --csv --log-file $OUT/launches_edit.csv python tools/edit_timing.py 16 3 > $OUT/edit_under_ncu.log 2>&1
python tools/summarize_launches.py $OUT/launches_edit.csv 2>&1 | head -30
