#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k chain 2>&1 | tail -8
for pat in chain2 chain3 chain4 chain5 chain6; do
timeout 300 python scripts/kbench.py --n 16384 --rows 32,64,128 --iters 20 --pattern $pat
done
timeout 300 python scripts/kbench.py --n 4096 --rows 32,64,128 --iters 50 --pattern chain4
} 2>&1 | tee gpurun_out/chain.log
