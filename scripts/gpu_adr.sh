#!/bin/bash
# adr bring-up on the GPU: parity tests + a first timing of BASELINE configs[3]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_adr_gpu.py -m gpu -x -q 2>&1 | tail -25
echo "=== config 4 timing: 2048^2 Strang RKC fixed_h 1e-3 tf 0.05"
( cd /tmp && B200_STATS=1 timeout 600 $OLDPWD/ceda-demonstrations_b200/bin/adr2d_b200 --nx 2048 --ny 2048 --integrator 3 --sts_method 0 --fixed_h 1e-3 --tf 0.05 --nout 1 --output 0 ; echo rc=$? )
( cd /tmp && B200_STATS=1 timeout 600 $OLDPWD/ceda-demonstrations_b200/bin/adr2d_b200 --nx 2048 --ny 2048 --integrator 3 --sts_method 1 --fixed_h 1e-3 --tf 0.05 --nout 1 --output 1 | tail -60; ls -la /tmp/solution.dat )
} 2>&1 | tee gpurun_out/adr.log
