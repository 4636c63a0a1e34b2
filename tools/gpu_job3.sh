#!/bin/bash
OUT=gpurun_out/j3; mkdir -p $OUT /tmp/mcmc
timeout 900 python -m pytest tests/test_gpu_delta.py tests/test_gpu_mcmc.py -m gpu -x -q > $OUT/pytest.log 2>&1; echo "exit $?" >> $OUT/pytest.log; tail -25 $OUT/pytest.log
python - > $OUT/gen.log 2>&1 <<'PY'
import delphy_b200 as db
from delphy_b200.maple import write_maple
emat, sites, info = db.synth_generate(db.synth_params(3))
write_maple(emat, sites, "/tmp/mcmc/t10k.maple", info["t_max_tip"])
PY
DPHY_DROPIN_STATS=1 timeout 300 delphy_b200/adapter/_build/delphy_b200_cli --v0-in-maple /tmp/mcmc/t10k.maple --v0-steps 200000 --v0-threads 1 --v0-seed 5 --v0-log-every 100000 2>&1 | grep -E "drop-in|Step 200000" | cut -c1-200 > $OUT/calls.txt; cat $OUT/calls.txt
tools/gpu_round.sh j3n l > $OUT/round.log 2>&1; tail -40 gpurun_out/j3n/launches_summary.txt
