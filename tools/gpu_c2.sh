#!/bin/bash
# config 2 (1,600 tips x 18,959 sites, nu_l on, 256 chains): launch list + one full capture of the general log-G kernels
O=gpurun_out/c2; mkdir -p $O
A="--config 2 --chains 256 --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-partitioned --no-mcmc --evals-per-step 2 --spr-studies 0"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches.csv python bench.py $A > $O/under_ncu.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches_summary.txt 2>&1; grep -E "emat_log_G|kernel  " $O/launches_summary.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"emat_log_G_tile_kernel|emat_log_G_straddler_kernel|emat_log_G_tree_kernel" -s 9 -c 3 -o $O/prof python bench.py $A > $O/ncu_full.log 2>&1
ncu -i $O/prof.ncu-rep --page raw --csv > $O/ncu_raw.csv 2>/dev/null
python tools/summarize_ncu.py $O/ncu_raw.csv > $O/ncu_summary.txt 2>&1
grep -E "Kernel Name|gpu__time_duration|dram__bytes_(read|write)|registers_per_thread|warps_active|l1tex__throughput|issue_active|long_scoreboard|stalled_barrier|inst_executed.sum|thread_inst_executed_per" $O/ncu_summary.txt
ncu -i $O/prof.ncu-rep --page source --csv -k regex:emat_log_G_tile_kernel > $O/tile_source.csv 2>/dev/null; ls -la $O
