#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
python __graft_entry__.py smoke 2>&1 | tail -2
python scripts/kbench.py --n 16384 --rows 8,16,32,64,128 --iters 30
python scripts/kbench.py --n 4096 --rows 8,16,32,64 --iters 100
python scripts/kbench.py --n 16384 --rows 32 --pattern rhs
python scripts/kbench.py --n 16384 --rows 32 --pattern final
echo "=== parity regressions"
python scripts/compare_runs.py -- --nx 128 --ny 128 --integrator rkc --tf 1 --nout 1 | tail -3
python scripts/compare_runs.py -- --nx 256 --ny 192 --integrator rkl --kx 1 --ky 0.1 --inhomogeneous --fixedstep 0.0009765625 --tf 0.0078125 --nout 1 | tail -2
python scripts/compare_runs.py -- --nx 128 --ny 128 --integrator erk --order -4 --fixedstep 0.0001220703125 --tf 0.0009765625 --nout 1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu-baseline
} 2>&1 | tee gpurun_out/iter.log
