#!/bin/bash
# round 2, call Q (4 GPUs): the 2- and 4-rank parity tests again with everything of the second half of the round in
# (stage-1 head chains over deep halos, provenance, next-step error weights, library from several translation units),
# N = 4 (2 x 2 blocks) and N = 1 bench lines on the same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_gpu or four_gpu" --durations=5 2>&1 | tail -12 > $O/r2q_pytest_multigpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29546"
timeout 400 $TR bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e > $O/r2q_scale_n4.json 2> $O/r2q_scale_n4.err
timeout 400 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/r2q_scale_n1.json 2> $O/r2q_scale_n1.err
ls -la $O | tail -4
