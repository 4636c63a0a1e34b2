#!/bin/bash
OUT=gpurun_out/${1:-j4}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_spr.py tests/test_gpu_full_size.py tests/test_gpu_delta.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "exit $?" >> $OUT/pytest.log; tail -25 $OUT/pytest.log
timeout 600 python bench.py --no-secondary --no-partitioned --no-mcmc --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; tail -2 $OUT/bench.err
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","spr_candidates_per_s","spr_ms_per_batch","spr_regions_per_batch","gpu_launches_spr")}, d.get("roofline_spr"))
PY
