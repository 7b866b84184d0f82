#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "deep_halo or temporal or golden" 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench3.json | cut -c1-250
B200_CHAIN=3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-250
B200_CHAIN=5 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-250
B200_CHAIN=6 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-250
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --local-n 4096 --method rkl 2>/dev/null | cut -c1-250
B200_CHAIN=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --local-n 4096 --method rkl 2>/dev/null | cut -c1-250
} 2>&1 | tee gpurun_out/quick.log
