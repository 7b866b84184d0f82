#!/bin/bash
# round 2, call F (2 GPUs): peer exchange after the fence fix -- N=2 bench (peer / NCCL), per-kernel trace of a 2-rank run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
for mode in peer nccl; do
  if [ $mode = nccl ]; then export B200_HALO_NCCL=1; else unset B200_HALO_NCCL; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > $O/r2f_scale_n2_$mode.json 2> $O/r2f_scale_n2_$mode.err
done
unset B200_HALO_NCCL
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > $O/r2f_scale_n1.json 2> $O/r2f_scale_n1.err
python - > $O/r2f_trace_2rank.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, "tests")
import compare_runs as cr
args = "--nx 32768 --ny 16384 --integrator rkc --fixedstep 2.5e-5 --tf 5e-5 --nout 1 --output 0".split()
for env in ({}, {"B200_HALO_NCCL": "1"}):
    e = dict(env, B200_TRACE_LAUNCHES="1")
    procs_out = cr.run(cr.B200_BIN, args, 2, env_extra=e, timeout=300)
    print("=== env", env)
    print(procs_out[1][-6000:])
PY
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_gpu" 2>&1 | tail -4 > $O/r2f_pytest_2gpu.log
ls -la $O | tail -6
