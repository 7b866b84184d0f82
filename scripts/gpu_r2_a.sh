#!/bin/bash
# round 2, call A: the new full-geometry parity tests + first GPU run of k_adr_chain, baseline bench line,
# current timings of the BASELINE configs through the drop-in driver binaries (state resident, reference CLI)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nproc > $O/r2a_host.log; nvidia-smi -L >> $O/r2a_host.log; df -h /tmp | tail -1 >> $O/r2a_host.log; free -g | head -2 >> $O/r2a_host.log
python __graft_entry__.py smoke > $O/r2a_smoke.log 2>&1
timeout 1500 python -m pytest tests/test_adr_gpu.py -m gpu -x -q 2>&1 | tail -8 > $O/r2a_pytest_adr.log
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 --deselect tests/test_adr_gpu.py 2>&1 | tail -40 > $O/r2a_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > $O/r2a_bench_default.json 2> $O/r2a_bench_default.err
D=./ceda-demonstrations_b200/bin/diffusion_2D_b200
A=$PWD/ceda-demonstrations_b200/bin/adr2d_b200
{
echo "=== C1 128^2 rkc tf=1"
timeout 300 $D --nx 128 --ny 128 --integrator rkc --tf 1 --nout 1 --output 1 | grep -E "Total simulation|Steps|RHS fn evals|stages used|B200"
echo "=== C2 4096^2 rkl aniso inhomogeneous internaleig tf=1e-3"
timeout 600 $D --nx 4096 --ny 4096 --integrator rkl --kx 1 --ky 0.1 --inhomogeneous --internaleig --tf 1e-3 --nout 1 --output 1 | grep -E "Total simulation|Steps|RHS fn evals|stages used|DEE|B200|dom_eig"
echo "=== C3 16384^2 rkc adaptive tf=2e-3"
timeout 600 $D --nx 16384 --ny 16384 --integrator rkc --tf 2e-3 --nout 1 --output 1 | grep -E "Total simulation|Steps|Step attempts|Error test|RHS fn evals|stages used|B200"
echo "=== C5 8192^2 dirk order 3 pcg jacobi tf=1e-3"
timeout 900 $D --nx 8192 --ny 8192 --integrator dirk --order 3 --tf 1e-3 --nout 1 --output 1 | grep -E "Total simulation|Steps|RHS fn evals|LS iters|NLS iters|Prec|B200"
for chain in 1 2 4 6; do
echo "=== C4 adr 2048^2 strang rkc sts_chain $chain"
( cd /tmp && B200_STATS=1 timeout 600 $A --nx 2048 --ny 2048 --integrator 3 --sts_method 0 --fixed_h 1e-3 --tf 0.05 --nout 1 --output 0 --sts_chain $chain | tail -3 )
done
echo "=== C4 adr 2048^2 strang rkl"
( cd /tmp && B200_STATS=1 timeout 600 $A --nx 2048 --ny 2048 --integrator 3 --sts_method 1 --fixed_h 1e-3 --tf 0.05 --nout 1 --output 0 | tail -3 )
} > $O/r2a_configs.log 2>&1
ls -la $O | tail -8
