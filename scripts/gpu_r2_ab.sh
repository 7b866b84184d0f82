#!/bin/bash
# round 2, call AB (2 GPUs): final state with the BULK ring as the default -- smoke(), the full GPU suite (the 2-rank
# tests included), the headline bench as the driver runs it at N = 1 and N = 2, the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2ab_smoke.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -14 > $O/r2ab_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/r2ab_bench_c3.json 2> $O/r2ab_bench_c3.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29548"
timeout 400 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/r2ab_scale_n2.json 2> $O/r2ab_scale_n2.err
ls -la $O | grep r2ab
