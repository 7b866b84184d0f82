#!/bin/bash
# round 2, call U (1 GPU): the steady state of k_chain_march without wrap-around logic (733 instead of 900 instructions per
# three rows in the plain flavour) -- parity tests, then plain PF=3 / plain PF=4 / BULK PF=3 / BULK PF=4 alternating;
# ncu --set full of the plain PF=3 body kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2u_smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "bulk or temporal_blocking or chain or kernel_geometry" 2>&1 | tail -4 > $O/r2u_pytest_chain.log
for i in 1 2; do
  B200_CHAIN_BULK=0 B200_CHAIN_PF=3 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2u_c3_plain_pf3_$i.json 2> $O/r2u_c3_plain_pf3_$i.err
  B200_CHAIN_BULK=0 B200_CHAIN_PF=4 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2u_c3_plain_pf4_$i.json 2> $O/r2u_c3_plain_pf4_$i.err
  B200_CHAIN_BULK=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2u_c3_bulk_pf3_$i.json 2> $O/r2u_c3_bulk_pf3_$i.err
  B200_CHAIN_BULK=2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2u_c3_bulk_pf4_$i.json 2> $O/r2u_c3_bulk_pf4_$i.err
done
B200_CHAIN_BULK=0 B200_CHAIN_PF=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_march --launch-skip 30 --launch-count 1 \
  -o $O/r2u_chain4_plain_pf3_body -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2u_ncu_plain.log 2>&1
B200_CHAIN_BULK=0 B200_CHAIN_PF=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_march --launch-skip 30 --launch-count 1 \
  -o $O/r2u_chain4_plain_pf4_body -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2u_ncu_plain4.log 2>&1
ls -la $O | grep r2u_
