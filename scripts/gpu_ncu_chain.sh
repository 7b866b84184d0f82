#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_march -s 6 -c 2 -o gpurun_out/chain4_r01 python scripts/kbench.py --n 16384 --rows 128 --iters 4 --pattern chain4 > gpurun_out/ncu_chain.log 2>&1
tail -3 gpurun_out/ncu_chain.log
