#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1 | tee $O/g_pytest_gpu.log
timeout 300 python scripts/kbench.py --n 16384 --iters 12 --rows 64,128 --pattern chain4 --variant 0 --arith exact,fma --uniform 0,1 2>&1 | grep pattern | tee $O/g_kbench.log
timeout 300 python scripts/kbench.py --n 16384 --iters 12 --rows 128 --pattern chain5,chain6 --variant 0 --arith exact,fma --uniform 1 2>&1 | grep pattern | tee -a $O/g_kbench.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/g_bench_default.json 2> $O/g_bench_default.err; cat $O/g_bench_default.json | cut -c1-330
for cfg in "fma 4 0" "fma 5 0" "fma 6 0" "exact 5 0"; do
  set -- $cfg
  timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --arith $1 --chain $2 --chain-variant $3 > $O/g_bench_$1_k$2_v$3.json 2>> $O/g_bench_variants.err
  python - "$O/g_bench_$1_k$2_v$3.json" "$cfg" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], "value %.4e ms/step %.2f frac %.3f kernel %s clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"], d["clocks"]))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
done 2>&1 | tee $O/g_bench_variants.log
