#!/bin/bash
# multi-GPU pass: N = 1, 2 (and whatever --gpus allows) weak-scaling bench + 2-GPU parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs visible: $NG" | tee gpurun_out/multi.log
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline 2>> gpurun_out/multi.err | tee gpurun_out/scale_n1.json | cut -c1-400
for N in 2 4 8; do
  if [ $N -le $NG ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 2>> gpurun_out/multi.err | tee gpurun_out/scale_n$N.json | cut -c1-400
  fi
done
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k two_gpu 2>&1 | tail -5 | tee -a gpurun_out/multi.log
tail -20 gpurun_out/multi.err
