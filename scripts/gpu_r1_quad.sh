#!/bin/bash
# round-1, quad kernel + FMA flavour + pipelined e2e: tests, default bench, kernel sweep, bench variants, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
( time timeout 560 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) 2>&1 | tee $O/q_pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/q_bench_default.json 2> $O/q_bench_default.err
cat $O/q_bench_default.json
timeout 300 python scripts/kbench.py --n 16384 --iters 12 --rows 64,128 --pattern chain4 --variant 1,0 --arith exact,fma 2>&1 | grep pattern | tee $O/q_kbench.log
timeout 300 python scripts/kbench.py --n 16384 --iters 12 --rows 64,128 --pattern chain5,chain6 --variant 1 --arith exact,fma 2>&1 | grep pattern | tee -a $O/q_kbench.log
for cfg in "fma 4 1" "fma 5 1" "fma 6 1" "exact 5 1" "exact 4 0"; do
  set -- $cfg
  timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --arith $1 --chain $2 --chain-variant $3 > $O/q_bench_$1_k$2_v$3.json 2>> $O/q_bench_variants.err
  python - "$O/q_bench_$1_k$2_v$3.json" "$cfg" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], "value %.4e ms/step %.2f frac %.3f kernel %s clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"], d["clocks"]))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
done 2>&1 | tee $O/q_bench_variants.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/q_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/q_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_chain_quad -s 6 -c 2 -o $O/quad4_r01 python scripts/kbench.py --n 16384 --rows 64 --iters 4 --pattern chain4 > $O/q_ncu_full.log 2>&1
tail -2 $O/q_ncu_full.log
ls -la $O | head -30
