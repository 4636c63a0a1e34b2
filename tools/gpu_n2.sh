#!/bin/bash
# 2-GPU sanity of the bench as the driver launches it (without the MCMC leg)
OUT=gpurun_out/n2; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-mcmc > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench N=2 exit $?"; tail -3 $OUT/bench_n2.err
python - <<PY
import json
d=json.load(open("$OUT/bench_n2.json"))
print({k: d[k] for k in ("value","n_gpus","ms_per_step","spr_candidates_per_s")}, "e2e", d["e2e"]["value"], "edit", d["e2e_edit"]["value"])
print("partitioned", {k: d["partitioned"][k] for k in ("parts","value","ms_per_cycle","allreduce_ms","collective","totals_match_whole_tree")})
print("wire", d["wire_format"]["load_ms_per_tree"], "spr frac", d["roofline_spr"]["frac"])
PY
