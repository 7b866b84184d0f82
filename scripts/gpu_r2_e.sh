#!/bin/bash
# round 2, call E (1 GPU): is the headline line stable?  bench first, kernel microbenchmark, small-grid timings, bench again
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2e_bench_first.json 2> $O/r2e_bench_first.err
python scripts/kbench.py --n 16384 --iters 5 --rows 128 --pattern chain4 --variant 0 --uniform 1 > $O/r2e_kbench.log 2>&1
D=./ceda-demonstrations_b200/bin/diffusion_2D_b200
{
for n in 32 64 128 256; do
echo "=== ${n}^2 rkc tf=1 (run twice: the second is warm)"
for rep in 1 2; do timeout 300 $D --nx $n --ny $n --integrator rkc --tf 1 --nout 1 --output 1 | grep -E "Total simulation|^Steps|RHS fn evals|B200 kernel launches"; done
echo "--- reference, 1 rank"
MPISHIM_NP=1 ./oracle/_ref/diffusion_2D_ref --nx $n --ny $n --integrator rkc --tf 1 --nout 1 --output 1 | grep -E "Total simulation|^Steps|RHS fn evals"
done
} > $O/r2e_small_grids.log 2>&1
B200_TRACE_LAUNCHES=1 timeout 300 $D --nx 128 --ny 128 --integrator rkc --tf 1 --nout 1 --output 1 > $O/r2e_c1_trace.log 2>&1
timeout 600 python -m pytest tests/test_adr_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "adr or golden" 2>&1 | tail -5 > $O/r2e_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r2e_bench_last.json 2> $O/r2e_bench_last.err
ls -la $O | tail -6
