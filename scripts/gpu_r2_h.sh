#!/bin/bash
# round 2, call H (1 GPU): stage 1 folded into the first chain launch of a step (HEAD flavour of k_chain_march):
# full GPU suite, headline bench with and without it, launch list and ncu --set full of a body and a head launch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -14 > $O/r2h_pytest_gpu.log
python bench.py --steps 10 --warmup 3 > $O/r2h_bench_head.json 2> $O/r2h_bench_head.err
B200_NO_CHAIN_HEAD=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2h_bench_nohead.json 2> $O/r2h_bench_nohead.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2h_bench_head_again.json 2> $O/r2h_bench_head_again.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2h_launches_bench_16384_rkc_head.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2h_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain_march --launch-skip 45 --launch-count 2 \
  -o $O/r2h_chain4_head_and_body -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2h_ncu_full.log 2>&1
ls -la $O | tail -8
