#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k two_gpu 2>&1 | tail -15 | tee gpurun_out/multi2.log
