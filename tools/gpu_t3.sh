#!/bin/bash
mkdir -p gpurun_out/t3
for P in 1 0; do
DPHY_MAIN_STREAM_PRIORITY=$P timeout 300 python bench.py --no-secondary --no-partitioned --no-mcmc --no-cpu-baseline > gpurun_out/t3/bench_p$P.json 2> gpurun_out/t3/bench_p$P.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/t3/bench_p$P.json'))
print('prio $P: spr ms', d['spr_ms_per_batch'], 'frac', d['roofline_spr']['frac'], 'value', d['value'], 'e2e', d['e2e']['value'], 'edit', d['e2e_edit']['value'], 'gen', d['loglik_general_schedule']['frac'])"
done
