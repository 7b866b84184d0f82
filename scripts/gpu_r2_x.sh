#!/bin/bash
# round 2, call X (1 GPU): rows fitted to whole waves (rows_fit_waves) -- chain kernel at 2048^2, config c4 A/B, adr tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_adr_gpu.py tests/test_kernels_gpu.py -m gpu -q 2>&1 | tail -4 > $O/r2x_pytest_adr_kernels.log
for n in 2048 3072 1024; do
B200_NO_WAVE_FIT=1 python scripts/kbench.py --n $n --pattern chain4 --variant 0 --uniform 0 --rows 0 --iters 50 2>&1 | sed 's/^/nofit /' >> $O/r2x_kbench.log
python scripts/kbench.py --n $n --pattern chain4 --variant 0 --uniform 0 --rows 0 --iters 50 2>&1 | sed 's/^/fit   /' >> $O/r2x_kbench.log
done
for i in 1 2; do
B200_NO_WAVE_FIT=1 python bench.py --config c4 --no-cpu-baseline > $O/r2x_bench_c4_nofit_$i.json 2> $O/r2x_bench_c4_nofit_$i.err
python bench.py --config c4 --no-cpu-baseline > $O/r2x_bench_c4_fit_$i.json 2> $O/r2x_bench_c4_fit_$i.err
done
python bench.py --config c2 --no-cpu-baseline > $O/r2x_bench_c2_1.json 2> $O/r2x_bench_c2_1.err
python bench.py --config c2 --no-cpu-baseline > $O/r2x_bench_c2_2.json 2> $O/r2x_bench_c2_2.err
cat $O/r2x_kbench.log
