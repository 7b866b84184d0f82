#!/bin/bash
# 2 GPUs, short: deep-halo chain parity and one N=2 bench line with the final defaults (uniform flavour, automatic rows)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_gpu_temporal" 2>&1 | tail -3 | tee $O/l_pytest_2gpu.log
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 4 --warmup 3 --no-e2e 2>$O/l_n2.err | tee $O/l_scale_n2.json | cut -c1-260
