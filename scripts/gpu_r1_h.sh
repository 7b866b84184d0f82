#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
( time timeout 500 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "sweep" --durations=5 2>&1 | tail -25 ) 2>&1 | tee $O/h_pytest_sweep.log
