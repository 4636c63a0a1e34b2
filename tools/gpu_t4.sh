#!/bin/bash
mkdir -p gpurun_out/t4
timeout 600 python -m pytest tests/test_gpu_logg.py tests/test_gpu_full_size.py tests/test_gpu_partition.py -m gpu -x -q > gpurun_out/t4/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t4/pytest.log
tail -4 gpurun_out/t4/pytest.log
timeout 300 python bench.py --config 2 --chains 256 --spr-studies 0 --no-secondary --no-partitioned --no-mcmc --no-cpu-baseline > gpurun_out/t4/bench_c2.json 2> gpurun_out/t4/bench_c2.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/t4/bench_c2.json'))
print('cfg2: value', d['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'gen frac', d['loglik_general_schedule']['frac'], d['loglik_general_schedule']['launch_ms'])"
tail -3 gpurun_out/t4/bench_c2.err
