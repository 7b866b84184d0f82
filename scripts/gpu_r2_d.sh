#!/bin/bash
# round 2, call D (1 GPU): fused implicit path on hardware (GPU suite + config lines), host-time trace of the adr driver
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
A=$PWD/ceda-demonstrations_b200/bin/adr2d_b200
for chain in 1 4 6; do
( cd /tmp && B200_TRACE_LAUNCHES=1 B200_STATS=1 timeout 300 $A --nx 2048 --ny 2048 --integrator 3 --sts_method 0 --fixed_h 1e-3 --tf 0.01 --nout 1 --output 0 --sts_chain $chain > /root/repo/$O/r2d_adr_trace_chain$chain.log 2>&1 )
( cd /tmp && B200_STATS=1 timeout 300 $A --nx 2048 --ny 2048 --integrator 3 --sts_method 0 --fixed_h 1e-3 --tf 0.05 --nout 1 --output 0 --sts_chain $chain 2>&1 | tail -1 >> /root/repo/$O/r2d_adr_trace_chain$chain.log )
done
timeout 2400 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -20 > $O/r2d_pytest_gpu.log
for cfg in c5 c2 c4; do
  timeout 900 python bench.py --config $cfg > $O/r2d_bench_$cfg.json 2> $O/r2d_bench_$cfg.err
done
timeout 600 python bench.py --config c4 --no-cpu-baseline --config-args --sts_chain 4 > $O/r2d_bench_c4_chain4.json 2> $O/r2d_bench_c4_chain4.err
B200_NO_DQ_FUSION=1 timeout 600 python bench.py --config c5 --no-cpu-baseline --no-e2e --config-args --no-fusion > $O/r2d_bench_c5_unfused.json 2> $O/r2d_bench_c5_unfused.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2d_bench_default.json 2> $O/r2d_bench_default.err
ls -la $O | tail -12
