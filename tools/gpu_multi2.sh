#!/bin/bash
# 2-GPU sanity: the NCCL partition test, then the bench exactly as the driver launches it.  usage: tools/gpu_multi2.sh <tag> <N>
TAG=${1:-m2}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 600 python -m pytest tests/test_gpu_partition.py -m gpu -q > $OUT/pytest_partition.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_partition.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench N=$N exit $?"; tail -3 $OUT/bench_n$N.err
python - <<PY
import json
d=json.load(open("$OUT/bench_n$N.json"))
print({k: d[k] for k in ("value","n_gpus","ms_per_step","spr_candidates_per_s")}, "e2e", d["e2e"]["value"], "edit", d["e2e_edit"]["value"])
print("partitioned", {k: d["partitioned"][k] for k in ("parts","value","ms_per_cycle","allreduce_ms","collective","totals_match_whole_tree")})
print("mcmc", {k:(v["steps_per_s"] if isinstance(v,dict) else v) for k,v in d["mcmc"].items()})
for k,v in d.get("configs",{}).items(): print("cfg",k,round(v["value"]), v.get("spr_studies_per_batch"), round(v["spr_candidates_per_s"]/1e9,2),"G", round(v["spr_roofline_frac"],4))
PY
