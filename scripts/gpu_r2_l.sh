#!/bin/bash
# round 2, call L (1 GPU): full GPU suite (implicit-reaction fixtures, block solver), the five bench configurations with
# the final code, the ERK / DIRK rows of the reference's evaluation matrix through the driver's main() in ONE process
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -25 > $O/r2l_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $O/r2l_bench_c3.json 2> $O/r2l_bench_c3.err
python bench.py --config c2 > $O/r2l_bench_c2.json 2> $O/r2l_bench_c2.err
python bench.py --config c4 > $O/r2l_bench_c4.json 2> $O/r2l_bench_c4.err
python bench.py --config c5 > $O/r2l_bench_c5.json 2> $O/r2l_bench_c5.err
timeout 1500 python scripts/runtests_diffusion2d_b200.py --inprocess --solvers dirk2-Jacobi,dirk3-Jacobi,erk2,erk3,erk4 \
  --grids 32,64 --rtols 1e-2,1e-4,1e-6 --hdivs 4,16,64 --out $O/r2l_sweep_erk_dirk_b200 > $O/r2l_sweep.log 2>&1
./ceda-demonstrations_b200/bin/diffusion_2D_b200 --nx 128 --ny 128 --integrator rkc --tf 1 --nout 1 --output 1 > $O/r2l_c1_run.log 2>&1
ls -la $O | tail -8
