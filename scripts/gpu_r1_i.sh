#!/bin/bash
# F4: the reference's evaluation matrix (RKC / RKL rows) on the B200 driver; CPU reference beside it for 32^2 / 64^2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
( time timeout 420 python scripts/runtests_diffusion2d_b200.py --solvers rkc,rkl --series adaptive --out $O/i_sweep_rkc_rkl ) 2>&1 | tail -5
( time timeout 200 python scripts/runtests_diffusion2d_b200.py --solvers rkc,rkl --series fixed --grids 128 --out $O/i_sweep_rkc_rkl_128 ) 2>&1 | tail -5
( time timeout 300 python scripts/runtests_diffusion2d_b200.py --solvers rkc,rkl --series adaptive --grids 32,64 --cpu --out $O/i_sweep_vs_cpu ) 2>&1 | tail -5
