#!/bin/bash
OUT=gpurun_out/${1:-s3j}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_delta.py tests/test_gpu_logg.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "exit $?" >> $OUT/pytest.log; tail -5 $OUT/pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-secondary --no-partitioned --no-mcmc --spr-studies 0 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
for k in ("value","e2e","e2e_edit"): print(k, d.get(k))
PY
