#!/bin/bash
# first GPU contact: parity on a handful of configs + raw timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc
{
echo "=== C1 128^2 rkc adaptive"
python scripts/compare_runs.py -- --nx 128 --ny 128 --integrator rkc --tf 1 --nout 1
echo "=== rkl inhomogeneous fixed"
python scripts/compare_runs.py -- --nx 256 --ny 192 --integrator rkl --kx 1 --ky 0.1 --inhomogeneous --fixedstep 0.0009765625 --tf 0.0078125 --nout 1
echo "=== odd nx (generic kernel) rkc adaptive"
python scripts/compare_runs.py -- --nx 101 --ny 77 --integrator rkc --tf 0.1 --nout 2
echo "=== internaleig rkl inhomogeneous"
python scripts/compare_runs.py -- --nx 256 --ny 256 --integrator rkl --inhomogeneous --kx 1 --ky 0.1 --internaleig --tf 0.1 --nout 1
echo "=== ssp104"
python scripts/compare_runs.py -- --nx 128 --ny 128 --integrator erk --order -4 --tf 0.05 --nout 1
echo "=== dirk pcg"
python scripts/compare_runs.py -- --nx 256 --ny 256 --integrator dirk --order 3 --tf 0.1 --nout 1
echo "=== timing 4096 rkc fixed"
./ceda-demonstrations_b200/bin/diffusion_2D_b200 --nx 4096 --ny 4096 --integrator rkc --fixedstep 1e-4 --tf 1e-3 --nout 1 --output 1
echo "=== timing 16384 rkc fixed"
./ceda-demonstrations_b200/bin/diffusion_2D_b200 --nx 16384 --ny 16384 --integrator rkc --fixedstep 1e-4 --tf 3e-4 --nout 1 --output 1
} 2>&1 | tee gpurun_out/first.log
