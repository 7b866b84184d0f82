#!/bin/bash
# round 2, call Z (1 GPU): uniform flavour with the x-direction products shared between neighbours (408 instead of 432
# FP64 instructions per three rows) -- chain / uniform / geometry parity tests, headline bench twice, ncu of the body kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "chain or uniform or bulk or temporal or geometry or baseline_size or nvector" 2>&1 | tail -4 > $O/r2z_pytest_chain.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2z_c3_1.json 2> $O/r2z_c3_1.err
B200_CHAIN_BULK=2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2z_c3_bulk_pf4_1.json 2> $O/r2z_c3_bulk_pf4_1.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2z_c3_2.json 2> $O/r2z_c3_2.err
B200_CHAIN_BULK=2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2z_c3_bulk_pf4_2.json 2> $O/r2z_c3_bulk_pf4_2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_march --launch-skip 30 --launch-count 1 \
  -o $O/r2z_chain4_body -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2z_ncu.log 2>&1
ls -la $O | grep r2z_
