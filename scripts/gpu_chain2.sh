#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -15
for k in 1 2 3 4 6; do
echo "=== bench chain=$k"
B200_CHAIN=$k timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-330
done
} 2>&1 | tee gpurun_out/chain2.log
