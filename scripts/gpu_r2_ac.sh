#!/bin/bash
# round 2, call AC (2 GPUs): N = 2 with the plain ring against the BULK ring (the deep-halo flavour), alternating
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29549"
for v in 0 2 0 2; do
B200_CHAIN_BULK=$v timeout 200 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e >> $O/r2ac_n2_bulk$v.json 2>> $O/r2ac_n2.err
done
