#!/bin/bash
# round 2, call C (1 GPU): full GPU suite with the constant-Value / peer-halo changes, bench line, launch lists of the
# adr driver with and without temporal blocking, single-GPU cost of the deep-halo exchange flavours
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > $O/r2c_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > $O/r2c_bench_default.json 2> $O/r2c_bench_default.err
D=./ceda-demonstrations_b200/bin/diffusion_2D_b200
A=$PWD/ceda-demonstrations_b200/bin/adr2d_b200
{
for extra in "" "--force-halo" "--force-halo --halo-nccl"; do
echo "=== 16384^2 rkc fixed 1e-4, 5 steps, $extra"
timeout 300 $D --nx 16384 --ny 16384 --integrator rkc --fixedstep 1e-4 --tf 5e-4 --nout 1 --output 0 $extra | grep -E "Total simulation|Steps|RHS fn evals|B200 kernel launches|chained"
done
} > $O/r2c_halo_flavours.log 2>&1
for chain in 1 4 6; do
( cd /tmp && timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file /root/repo/$O/r2c_adr_launches_chain$chain.csv $A --nx 2048 --ny 2048 --integrator 3 --sts_method 0 --fixed_h 1e-3 --tf 0.003 --nout 1 --output 0 --sts_chain $chain > /dev/null 2>&1 )
done
ls -la $O | tail -8
