#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) 2>&1 | tee $O/j_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/j_smoke.log
timeout 400 python bench.py > $O/j_bench_default.json 2> $O/j_bench_default.err; cut -c1-300 $O/j_bench_default.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/j_bench_reference.json 2>> $O/j_bench_default.err; cat $O/j_bench_reference.json | cut -c1-600
for rows in 128 32; do
  timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --chain-rows $rows > $O/j_bench_rows$rows.json 2>> $O/j_bench_default.err
  python - "$O/j_bench_rows$rows.json" "rows $rows" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], "value %.4e ms/step %.2f frac %.3f clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"]))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
done 2>&1 | tee $O/j_bench_rows.log
