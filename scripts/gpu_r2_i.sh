#!/bin/bash
# round 2, call I (1 GPU): full GPU suite with the HEAD chain and the speculative next-step error weights; small grids
# (32^2..256^2 adaptive RKC, BASELINE configs[0] = 128^2) against the 1-core reference, with the speculation on / off;
# launch trace of C1; config c2 bench (adaptive, large) as a regression check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -25 > $O/r2i_pytest_gpu.log
D=./ceda-demonstrations_b200/bin/diffusion_2D_b200
{
for n in 32 64 128 256; do
echo "=== ${n}^2 rkc tf=1 (run three times: the later ones are warm)"
for rep in 1 2 3; do timeout 300 $D --nx $n --ny $n --integrator rkc --tf 1 --nout 1 --output 1 | grep -E "Total simulation|^Steps|RHS fn evals|B200 kernel launches"; done
echo "--- same, B200_NO_SPEC_EWT=1"
for rep in 1 2; do B200_NO_SPEC_EWT=1 timeout 300 $D --nx $n --ny $n --integrator rkc --tf 1 --nout 1 --output 1 | grep -E "Total simulation|B200 kernel launches"; done
echo "--- reference, 1 rank"
for rep in 1 2; do MPISHIM_NP=1 ./oracle/_ref/diffusion_2D_ref --nx $n --ny $n --integrator rkc --tf 1 --nout 1 --output 1 | grep -E "Total simulation|^Steps|RHS fn evals"; done
done
} > $O/r2i_small_grids.log 2>&1
B200_TRACE_LAUNCHES=1 timeout 300 $D --nx 128 --ny 128 --integrator rkc --tf 1 --nout 1 --output 1 > $O/r2i_c1_trace.log 2>&1
B200_TRACE_LAUNCHES=1 timeout 300 $D --nx 32 --ny 32 --integrator rkc --tf 1 --nout 1 --output 1 > $O/r2i_32_trace.log 2>&1
python bench.py --config c2 > $O/r2i_bench_c2.json 2> $O/r2i_bench_c2.err
ls -la $O | tail -6
