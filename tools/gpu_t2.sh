#!/bin/bash
mkdir -p gpurun_out/t2
timeout 600 python -m pytest tests/test_gpu_spr.py tests/test_gpu_full_size.py tests/test_gpu_mcmc.py -m gpu -x -q > gpurun_out/t2/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t2/pytest.log
tail -5 gpurun_out/t2/pytest.log
timeout 300 python bench.py --no-secondary --no-partitioned --no-mcmc --no-cpu-baseline > gpurun_out/t2/bench.json 2> gpurun_out/t2/bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/t2/bench.json'))
print('spr ms', d['spr_ms_per_batch'], 'frac', d['roofline_spr']['frac'], 'value', d['value'])"
DPHY_SPR_TAIL_STREAM=1 timeout 300 python bench.py --no-secondary --no-partitioned --no-mcmc --no-cpu-baseline > gpurun_out/t2/bench_notail.json 2> gpurun_out/t2/bench_notail.err
python -c "
import json; d=json.load(open('gpurun_out/t2/bench_notail.json'))
print('tail mode 1: spr ms', d['spr_ms_per_batch'], 'frac', d['roofline_spr']['frac'])"
