#!/bin/bash
# round 2, call Y (2 GPUs): the 2-rank parity tests and the N = 2 / N = 1 bench lines with the final kernels (wrap-free
# steady state in the deep-halo flavour, rows fitted to waves)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpu" --durations=5 2>&1 | tail -12 > $O/r2y_pytest_multigpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547"
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/r2y_scale_n2.json 2> $O/r2y_scale_n2.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/r2y_scale_n1.json 2> $O/r2y_scale_n1.err
ls -la $O | grep r2y
