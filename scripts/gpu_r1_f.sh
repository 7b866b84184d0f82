#!/bin/bash
# 2 GPUs: multi-rank parity tests, N=1 and N=2 bench lines (pipelined e2e on both ranks)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
{
timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_gpu" 2>&1 | tail -6
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline 2>$O/f_n1.err | tee $O/f_scale_n1.json | cut -c1-400
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 2>$O/f_n2.err | tee $O/f_scale_n2.json | cut -c1-400
tail -5 $O/f_n2.err
} 2>&1 | tee $O/f_multi2.log
