#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) 2>&1 | tee $O/k_pytest_gpu.log
timeout 400 python bench.py > $O/k_bench_default.json 2> $O/k_bench_default.err; cut -c1-260 $O/k_bench_default.json
for rows in 64 192 256; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --chain-rows $rows > $O/k_bench_rows$rows.json 2>> $O/k_bench_default.err
  python - "$O/k_bench_rows$rows.json" "rows $rows" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], "value %.4e ms/step %.2f frac %.3f clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
done 2>&1 | tee $O/k_bench_rows.log
timeout 120 python scripts/kbench.py --n 4096 --iters 40 --rows 32,64,128 --pattern chain4 --variant 0 --uniform 1 2>&1 | grep pattern | tee $O/k_kbench_4096.log
