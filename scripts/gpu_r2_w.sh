#!/bin/bash
# round 2, call W (1 GPU): rows per block of the chain kernel against a wave-count model (kernel microbenchmark)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
for u in 0 1; do
python scripts/kbench.py --n 4096 --pattern chain4 --variant 0 --uniform $u --rows 0,32,35,40,47,59,70,94,141,147 --iters 50 >> $O/r2w_kbench_rows.log 2>&1
python scripts/kbench.py --n 8192 --pattern chain4 --variant 0 --uniform $u --rows 0,64,128,138,182,256,273 --iters 30 >> $O/r2w_kbench_rows.log 2>&1
python scripts/kbench.py --n 16384 --pattern chain4 --variant 0 --uniform $u --rows 0,222,256,293 --iters 20 >> $O/r2w_kbench_rows.log 2>&1
done
python scripts/kbench.py --n 4096 --pattern chain2,chain3,chain6 --variant 0 --uniform 0 --rows 0,32,35,47,70 --iters 50 >> $O/r2w_kbench_rows.log 2>&1
python scripts/kbench.py --n 2048 --pattern chain4 --variant 0 --uniform 0 --rows 0,16,19,24,32,37 --iters 50 >> $O/r2w_kbench_rows.log 2>&1
tail -5 $O/r2w_kbench_rows.log
