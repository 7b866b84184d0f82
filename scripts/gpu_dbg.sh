#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for ord in -2 -3 -4; do
echo "=== ssp order $ord fixed"
python scripts/compare_runs.py -- --nx 128 --ny 128 --integrator erk --order $ord --fixedstep 0.0001220703125 --tf 0.0009765625 --nout 1
done
echo "=== ssp104 adaptive no-fusion"
python scripts/compare_runs.py -- --nx 128 --ny 128 --integrator erk --order -4 --tf 0.05 --nout 1 --no-fusion 2>&1 | grep -v Unknown
echo "=== ssp s3 adaptive"
python scripts/compare_runs.py -- --nx 128 --ny 128 --integrator erk --order -3 --tf 0.05 --nout 1
echo "=== ssp s2 adaptive"
python scripts/compare_runs.py -- --nx 128 --ny 128 --integrator erk --order -2 --tf 0.05 --nout 1
echo "=== erk order 3 adaptive"
python scripts/compare_runs.py -- --nx 128 --ny 128 --integrator erk --order 3 --tf 0.05 --nout 1
} 2>&1 | tee gpurun_out/dbg.log
