#!/bin/bash
# round-1 iteration: trimmed chain kernel + adr marching kernel -- tests, kernel sweep, configs, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/it2_pytest_gpu.log
{
for pat in chain4; do timeout 300 python scripts/kbench.py --n 16384 --rows 32,64,128,256,512 --iters 20 --pattern $pat; done
for pat in chain2 chain3 chain5 chain6; do timeout 300 python scripts/kbench.py --n 16384 --rows 128,512 --iters 20 --pattern $pat; done
timeout 300 python scripts/kbench.py --n 4096 --rows 32,64,128,256 --iters 50 --pattern chain4
} 2>&1 | grep pattern | tee gpurun_out/it2_kbench.log
D=./ceda-demonstrations_b200/bin/diffusion_2D_b200
A=$PWD/ceda-demonstrations_b200/bin/adr2d_b200
{
echo "=== C1 128^2 rkc tf=1"
timeout 300 $D --nx 128 --ny 128 --integrator rkc --tf 1 --nout 1 --output 1 | grep -E "Total simulation|Steps|RHS fn evals|stages used|B200"
echo "=== C2b 4096^2 rkl aniso fixed 1e-4 tf=1e-3"
timeout 600 $D --nx 4096 --ny 4096 --integrator rkl --kx 1 --ky 0.1 --fixedstep 1e-4 --tf 1e-3 --nout 1 --output 1 | grep -E "Total simulation|Steps|RHS fn evals|stages used|B200"
echo "=== C4 adr 2048^2 strang rkc"
( cd /tmp && B200_STATS=1 timeout 600 $A --nx 2048 --ny 2048 --integrator 3 --sts_method 0 --fixed_h 1e-3 --tf 0.05 --nout 1 --output 0 | tail -1 )
echo "=== C4 adr 2048^2 strang rkl"
( cd /tmp && B200_STATS=1 timeout 600 $A --nx 2048 --ny 2048 --integrator 3 --sts_method 1 --fixed_h 1e-3 --tf 0.05 --nout 1 --output 0 | tail -1 )
} 2>&1 | tee gpurun_out/it2_configs.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/it2_bench_n1.json 2> gpurun_out/it2_bench_n1.err
cat gpurun_out/it2_bench_n1.json
