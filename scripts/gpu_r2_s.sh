#!/bin/bash
# round 2, call S (1 GPU): BULK flavour of k_chain_march (cp.async.bulk + mbarrier operand ring) -- parity tests, A/B of
# the headline bench against the plain flavour (alternating), ncu --set full of the BULK body kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2s_smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "bulk" --durations=5 2>&1 | tail -14 > $O/r2s_pytest_bulk.log
for i in 1 2; do
  B200_CHAIN_BULK=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2s_c3_plain_$i.json 2> $O/r2s_c3_plain_$i.err
  B200_CHAIN_BULK=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2s_c3_bulk_$i.json 2> $O/r2s_c3_bulk_$i.err
done
B200_CHAIN_BULK=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_march --launch-skip 30 --launch-count 1 \
  -o $O/r2s_chain4_bulk_body -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2s_ncu_bulk.log 2>&1
ls -la $O | grep r2s_
