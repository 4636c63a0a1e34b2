#!/bin/bash
# One GPU round: parity tests, bench, ncu launch list, ncu full captures.  usage: tools/gpu_round.sh <tag> [sections]
# sections: any of  t (pytest -m gpu)  b (bench)  l (launch list)  n (ncu --set full of the log-G + SPR kernels)  s (smoke)
TAG=${1:-r}; SEC=${2:-tbln}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
if [[ $SEC == *s* ]]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
  tail -3 $OUT/smoke.log
fi
if [[ $SEC == *t* ]]; then
  timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -15 $OUT/pytest_gpu.log
fi
if [[ $SEC == *b* ]]; then
  timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
  tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
fi
if [[ $SEC == *l* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-partitioned --no-mcmc --evals-per-step 2 --spr-batches-per-step 1 > $OUT/bench_under_ncu.log 2>&1
  python tools/summarize_launches.py $OUT/launches.csv > $OUT/launches_summary.txt 2>&1; cat $OUT/launches_summary.txt
fi
if [[ $SEC == *n* ]]; then
  # one full capture of each hot kernel (the 4th..: warm-up launches are skipped with -s)
  timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"emat_log_G_folded_kernel|emat_log_G_tile_kernel|spr_" -s 48 -c 44 \
    -o $OUT/prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-partitioned --no-mcmc --evals-per-step 2 --spr-batches-per-step 1 > $OUT/ncu_full.log 2>&1
  ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/ncu_raw.csv 2>/dev/null
  python tools/summarize_ncu.py $OUT/ncu_raw.csv > $OUT/ncu_summary.txt 2>&1; grep -E "Kernel Name|gpu__time_duration|dram__bytes_(read|write)|dram__bytes.sum.per" $OUT/ncu_summary.txt
fi
if [[ $SEC == *r* ]]; then
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_arm.json 2> $OUT/bench_ref.err; echo "ref exit $?"; cut -c1-300 $OUT/bench_reference_arm.json
fi
ls -la $OUT
