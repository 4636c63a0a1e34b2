#!/bin/bash
# round-2 GPU check: tests, bench N=1, optionally bench N=2 under torchrun.  usage: tools/gpu_r2.sh <tag> [t][b][2][r]
TAG=${1:-r2}; SEC=${2:-tb}
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [[ $SEC == *t* ]]; then
  timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -12 $OUT/pytest_gpu.log
fi
if [[ $SEC == *b* ]]; then
  timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -3 $OUT/bench.err; head -c 6000 $OUT/bench.json
fi
if [[ $SEC == *2* ]]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --no-secondary > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench2 exit $?"; tail -3 $OUT/bench_n2.err; head -c 5000 $OUT/bench_n2.json
fi
if [[ $SEC == *r* ]]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref exit $?"; head -c 3000 $OUT/bench_ref.json
fi
