#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python scripts/pipe_diag.py 16384 6 > $O/b_pipe_diag.log 2>&1; cat $O/b_pipe_diag.log | tail -60
timeout 300 python scripts/kbench.py --n 16384 --iters 12 --rows 64 --pattern chain4,chain3 --variant 2,0 --arith exact,fma 2>&1 | grep pattern | tee $O/b_kbench.log
timeout 300 python scripts/kbench.py --n 16384 --iters 12 --rows 64,128 --pattern chain5,chain6 --variant 0 --arith exact,fma 2>&1 | grep pattern | tee -a $O/b_kbench.log
( time timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "chain" 2>&1 | tail -5 ) 2>&1 | tee $O/b_pytest_chain.log
