#!/bin/bash
# round 2, call B (2 GPUs): peer-mapped deep-halo exchange against the NCCL one -- parity tests and the N=2 bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r2b_topo.log 2>&1
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_gpu or deep_halo" 2>&1 | tail -15 > $O/r2b_pytest_2gpu.log
B200_HALO_NCCL=1 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_gpu_temporal" 2>&1 | tail -5 > $O/r2b_pytest_2gpu_nccl.log
for mode in peer nccl; do
  if [ $mode = nccl ]; then export B200_HALO_NCCL=1; else unset B200_HALO_NCCL; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > $O/r2b_scale_n2_$mode.json 2> $O/r2b_scale_n2_$mode.err
done
unset B200_HALO_NCCL
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2b_scale_n1.json 2> $O/r2b_scale_n1.err
ls -la $O | tail -8
