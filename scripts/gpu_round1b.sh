#!/bin/bash
# measurement pass with temporal blocking on: bench, launch list, ncu full of the chain kernel, tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench2.err | tee gpurun_out/bench2.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_march -s 30 -c 2 -o gpurun_out/chain_r01b python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
