#!/bin/bash
mkdir -p gpurun_out/t5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t5/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/t5/smoke.log; tail -3 gpurun_out/t5/smoke.log
timeout 600 python -m pytest tests/test_gpu_spr.py tests/test_gpu_api_tree.py -m gpu -x -q > gpurun_out/t5/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/t5/pytest.log
tail -5 gpurun_out/t5/pytest.log
timeout 300 python bench.py --no-partitioned --no-mcmc --cpu-seconds 2 --spr-studies 0 > gpurun_out/t5/bench.json 2> gpurun_out/t5/bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/t5/bench.json'))
print('wire', d['wire_format']['load_ms_per_tree'], d['wire_format']['write_ms_per_tree'], 'cfg2', d['configs']['2']['value'], d['configs']['2']['roofline'])"
