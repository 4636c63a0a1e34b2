#!/bin/bash
# the wire-format tests on the GPU box (+ the loader's cost on a 100k-tip tree)
mkdir -p gpurun_out/api
timeout 600 python -m pytest tests/test_gpu_api_tree.py -m gpu -x -q > gpurun_out/api/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/api/pytest.log
tail -30 gpurun_out/api/pytest.log
timeout 300 python tools/api_tree_timing.py > gpurun_out/api/timing.log 2>&1; tail -20 gpurun_out/api/timing.log
