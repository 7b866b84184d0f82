#!/bin/bash
# last GPU seconds of round 1: ncu of the FINAL default chain kernel (uniform flavour, 128 rows) + launch list of a bench step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 60 ncu --set full --clock-control none --import-source on -k regex:k_chain_march -s 6 -c 1 -o /tmp/march4_uni python scripts/kbench.py --n 16384 --iters 3 --rows 128 --pattern chain4 --variant 0 --uniform 1 > $O/m_ncu_march4_uni.log 2>&1
ncu -i /tmp/march4_uni.ncu-rep --page raw --csv > $O/m_march4_uni_raw.csv 2>/dev/null
ncu -i /tmp/march4_uni.ncu-rep --page details > $O/m_march4_uni_details.txt 2>/dev/null
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/m_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/m_ncu_launches.log 2>&1
ls -la $O | tail -6
