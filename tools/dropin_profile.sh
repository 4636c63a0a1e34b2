#!/bin/bash
# where the drop-in's time goes on a small tree: per-thread residency stats + wall clock.  usage: tools/dropin_profile.sh <tag>
OUT=gpurun_out/${1:-dp}; mkdir -p $OUT /tmp/mcmc
python - > $OUT/gen.log 2>&1 <<'PY'
import delphy_b200 as db
from delphy_b200.maple import write_maple
emat, sites, info = db.synth_generate(db.synth_params(1))
write_maple(emat, sites, "/tmp/mcmc/cfg1.maple", info["t_max_tip"])
PY
for bin in oracle/_ref/delphy delphy_b200/adapter/_build/delphy_b200_cli; do
  s=$(date +%s.%N)
  DPHY_DROPIN_STATS=1 $bin --v0-in-maple /tmp/mcmc/cfg1.maple --v0-steps 2000000 --v0-threads 1 --v0-seed 2 --v0-log-every 1000000 2>&1 | grep -E "drop-in|Step 2000000|rror" | cut -c1-260
  e=$(date +%s.%N); echo "$bin wall $(echo "$e - $s" | bc) s"
done > $OUT/profile.txt 2>&1
cat $OUT/profile.txt $OUT/gen.log
