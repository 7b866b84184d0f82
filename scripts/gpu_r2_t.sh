#!/bin/bash
# round 2, call T (1 GPU): prefetch depth of the BULK flavour -- plain / BULK PF=3 / BULK PF=4 alternating, ncu of PF=4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "bulk" 2>&1 | tail -4 > $O/r2t_pytest_bulk.log

for i in 1 2; do
  for v in 0 1 2; do
    B200_CHAIN_BULK=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2t_c3_bulk${v}_$i.json 2> $O/r2t_c3_bulk${v}_$i.err
  done
done
B200_CHAIN_BULK=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_march --launch-skip 30 --launch-count 1 \
  -o $O/r2t_chain4_bulk_pf4_body -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2t_ncu_bulk.log 2>&1
ls -la $O | grep r2t_
