#!/bin/bash
# Round-1 measurement pass on one B200: smoke, bench, ncu launch list, ncu full capture of the
# fused stage kernel, kernel microbench sweep, GPU test suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv | tee gpurun_out/gpu.txt
nproc | tee -a gpurun_out/gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench1.err | tee gpurun_out/bench1.json | tail -3
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stage_march -s 100 -c 3 -o gpurun_out/stage_r01 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
{
timeout 300 python scripts/kbench.py --n 16384 --rows 4,8,16,32,64,128 --iters 30
timeout 300 python scripts/kbench.py --n 4096 --rows 4,8,16,32,64 --iters 100
timeout 300 python scripts/kbench.py --n 16384 --rows 8 --pattern rhs
timeout 300 python scripts/kbench.py --n 16384 --rows 8 --pattern final
} 2>&1 | tee gpurun_out/kbench.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
ls -la gpurun_out
