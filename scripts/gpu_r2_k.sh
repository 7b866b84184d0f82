#!/bin/bash
# round 2, call K (1 GPU): small grids after (a) next-step error weights out of the closing stage, (b) adaptive steps
# whose first chain starts from y_n alone (provenance), (c) waiting for the reduction VALUE instead of the stream;
# host profile (time inside launch calls / waits) to see what is left
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
D=./ceda-demonstrations_b200/bin/diffusion_2D_b200
{
for n in 32 64 128 256; do
echo "=== ${n}^2 rkc tf=1 (run three times: the later ones are warm)"
for rep in 1 2 3; do timeout 300 $D --nx $n --ny $n --integrator rkc --tf 1 --nout 1 --output 1 | grep -E "Total simulation|^Steps|RHS fn evals|B200 kernel launches"; done
echo "--- same, B200_NO_POLL=1"
for rep in 1 2; do B200_NO_POLL=1 timeout 300 $D --nx $n --ny $n --integrator rkc --tf 1 --nout 1 --output 1 | grep -E "Total simulation|B200 kernel launches"; done
echo "--- same, B200_HOST_PROFILE=1"
B200_HOST_PROFILE=1 timeout 300 $D --nx $n --ny $n --integrator rkc --tf 1 --nout 1 --output 1 2>&1 | grep -E "Total simulation|host profile"
echo "--- reference, 1 rank"
for rep in 1 2 3; do MPISHIM_NP=1 ./oracle/_ref/diffusion_2D_ref --nx $n --ny $n --integrator rkc --tf 1 --nout 1 --output 1 | grep -E "Total simulation"; done
done
} > $O/r2k_small_grids.log 2>&1
B200_TRACE_LAUNCHES=1 timeout 300 $D --nx 128 --ny 128 --integrator rkc --tf 1 --nout 1 --output 1 > $O/r2k_c1_trace.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -k "golden or sweep or head or error_weights or temporal or live or nvector" 2>&1 | tail -8 > $O/r2k_pytest_gpu.log
ls -la $O | tail -4
