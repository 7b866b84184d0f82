#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
( time timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -x -q -k "chain or fma or pipelined or variants" 2>&1 | tail -5 ) 2>&1 | tee $O/c_pytest_chain.log
timeout 300 python scripts/kbench.py --n 16384 --iters 12 --rows 64,128 --pattern chain4 --variant 1,2,0 --arith exact,fma 2>&1 | grep pattern | tee $O/c_kbench.log
timeout 300 python scripts/kbench.py --n 16384 --iters 12 --rows 64,128 --pattern chain5,chain6 --variant 1 --arith exact,fma 2>&1 | grep pattern | tee -a $O/c_kbench.log
timeout 300 python scripts/pipe_diag.py 16384 6 > $O/c_pipe_diag.log 2>&1; cat $O/c_pipe_diag.log | tail -40
