#!/bin/bash
# round 2, call N (1 GPU): final numbers with the chain kernels preloaded at session creation: full GPU suite, c2 (three
# times: it was erratic with lazy module loading), c3 headline (20 steps, as the driver's scaling run), c4, C1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -14 > $O/r2n_pytest_gpu.log
for i in 1 2 3; do python bench.py --config c2 --no-cpu-baseline --no-e2e > $O/r2n_bench_c2_$i.json 2> $O/r2n_bench_c2_$i.err; done
python bench.py --config c2 > $O/r2n_bench_c2.json 2> $O/r2n_bench_c2.err
python bench.py --steps 20 --warmup 5 > $O/r2n_bench_c3.json 2> $O/r2n_bench_c3.err
python bench.py --config c4 > $O/r2n_bench_c4.json 2> $O/r2n_bench_c4.err
D=./ceda-demonstrations_b200/bin/diffusion_2D_b200
for rep in 1 2 3; do $D --nx 128 --ny 128 --integrator rkc --tf 1 --nout 1 --output 1 | grep -E "Total simulation|B200 kernel launches"; done > $O/r2n_c1.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2n_bench_reference_arm.json 2> $O/r2n_bench_reference_arm.err
ls -la $O | tail -5
