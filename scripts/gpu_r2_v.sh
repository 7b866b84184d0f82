#!/bin/bash
# round 2, call V (1 GPU): state after the wrap-free steady state of k_chain_march / k_adr_chain -- smoke(), full GPU
# suite, the headline bench as the driver runs it, configs c2 / c4 / c5, launch list of the headline (kernel shares)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2v_smoke.log 2>&1
timeout 1200 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -14 > $O/r2v_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/r2v_bench_c3.json 2> $O/r2v_bench_c3.err
python bench.py --config c2 --no-cpu-baseline > $O/r2v_bench_c2.json 2> $O/r2v_bench_c2.err
python bench.py --config c4 --no-cpu-baseline > $O/r2v_bench_c4.json 2> $O/r2v_bench_c4.err
python bench.py --config c5 --no-cpu-baseline > $O/r2v_bench_c5.json 2> $O/r2v_bench_c5.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2v_launches_bench_16384_rkc.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2v_ncu_list.log 2>&1
ls -la $O | grep r2v_
