#!/bin/bash
# round 2, call R (1 GPU): final state -- smoke(), full GPU suite, headline bench twice (256 rows per block), launch lists
# of one config-5 step (the PCG iteration's kernels) and one config-4 step, ncu --set full of k_dq_march and k_adr_chain
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2r_smoke.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -14 > $O/r2r_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/r2r_bench_c3.json 2> $O/r2r_bench_c3.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2r_bench_c3_again.json 2> $O/r2r_bench_c3_again.err
B200_BENCH_PER_STEP=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2r_bench_c3_per_step.json 2> $O/r2r_bench_c3_per_step.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2r_launches_c5_one_step.csv \
  python bench.py --config c5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2r_ncu_c5_list.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2r_launches_c4_steps.csv \
  python bench.py --config c4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2r_ncu_c4_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dq_march --launch-skip 200 --launch-count 1 \
  -o $O/r2r_dq_march -f python bench.py --config c5 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2r_ncu_dq.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adr_chain --launch-skip 20 --launch-count 1 \
  -o $O/r2r_adr_chain -f python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2r_ncu_adr.log 2>&1
ls -la $O | tail -6
