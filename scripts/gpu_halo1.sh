#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -m gpu -x -q -k "deep_halo or temporal or chain" 2>&1 | tail -15
echo "=== bench wrap chain=4"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-330
echo "=== bench force-halo chain=4 (single rank, deep-halo kernel flavour)"
B200_FORCE_HALO=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-330
timeout 300 python scripts/kbench.py --n 16384 --rows 64,128 --iters 20 --pattern chain4
timeout 300 python scripts/kbench.py --n 16384 --rows 64,128 --iters 20 --pattern chain3
} 2>&1 | tee gpurun_out/halo1.log
