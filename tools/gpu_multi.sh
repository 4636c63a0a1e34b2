#!/bin/bash
# N-GPU bench exactly as the driver launches it, + the reference arm under torchrun.  usage: tools/gpu_multi.sh <tag> <N>
TAG=${1:-m}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "bench N=$N exit $?"; tail -3 $OUT/bench_n$N.err; cut -c1-900 $OUT/bench_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err; echo "ref N=$N exit $?"; cut -c1-300 $OUT/bench_ref_n$N.json
timeout 900 python -m pytest tests/test_gpu_full_size.py -m gpu -q > $OUT/pytest_full.log 2>&1; echo "pytest exit $?"; tail -12 $OUT/pytest_full.log
