#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 400 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_gpu" 2>&1 | tail -15
timeout 200 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tee gpurun_out/scale_chain_n1.json | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 2>/dev/null | tee gpurun_out/scale_chain_n2.json | cut -c1-300
} 2>&1 | tee gpurun_out/multi2.log
