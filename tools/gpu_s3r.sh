#!/bin/bash
timeout 120 python -m pytest tests/test_gpu_logg.py -m gpu -x -q 2>&1 | tail -3
timeout 90 python tools/tally_timing.py 16 4 2>&1 | tail -5; echo "tally exit $?"
