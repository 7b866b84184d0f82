#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1 | tee $O/d_pytest_gpu.log
timeout 300 python scripts/pipe_diag.py 16384 6 > $O/d_pipe_diag.log 2>&1; grep -v "b200_pipe\]" $O/d_pipe_diag.log | tail; grep "take\|put" $O/d_pipe_diag.log | tail -8
timeout 400 python bench.py --steps 5 --warmup 3 > $O/d_bench_default.json 2> $O/d_bench_default.err; cat $O/d_bench_default.json
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/d_bench_steps10.json 2>> $O/d_bench_default.err; cat $O/d_bench_steps10.json
for cfg in "exact 4 1" "fma 4 1" "fma 4 0" "fma 5 1"; do
  set -- $cfg
  timeout 200 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --arith $1 --chain $2 --chain-variant $3 > $O/d_bench_$1_k$2_v$3.json 2>> $O/d_bench_variants.err
  python - "$O/d_bench_$1_k$2_v$3.json" "$cfg" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], "value %.4e ms/step %.2f frac %.3f kernel %s clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"], d["clocks"]))
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
done 2>&1 | tee $O/d_bench_variants.log
