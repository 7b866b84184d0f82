#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/e_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/e_ncu_launches.log 2>&1
cap() { # name kernel-regex kbench-args...
  name=$1; rx=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s 6 -c 1 -o /tmp/$name python scripts/kbench.py --n 16384 --iters 3 "$@" > $O/e_ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/e_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details > $O/e_${name}_details.txt 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv --print-source sass > $O/e_${name}_source.csv 2>/dev/null
}
cap march4_final k_chain_march --rows 64 --pattern chain4 --variant 0
cap quad4_v2 k_chain_quad --rows 64 --pattern chain4 --variant 1
cap quad4_v2_fma k_chain_quad --rows 128 --pattern chain4 --variant 1 --arith fma
cap march4_fma k_chain_march --rows 128 --pattern chain4 --variant 0 --arith fma
du -sh $O; ls -la $O
