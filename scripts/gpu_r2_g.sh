#!/bin/bash
# round 2, call G (8 GPUs): multi-GPU parity at 2 / 4 / 8 ranks with the peer-mapped halos, host <-> device probe with all
# ranks copying at once (unbound / NUMA-bound), the N = 8 bench line (peer and NCCL halos), config 5 on 8 ranks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r2g_topo.log 2>&1
nproc >> $O/r2g_topo.log; cat /sys/devices/system/node/node*/cpulist >> $O/r2g_topo.log 2>&1; free -g | head -2 >> $O/r2g_topo.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "two_gpu or four_gpu or eight_gpu" --durations=6 2>&1 | tail -14 > $O/r2g_pytest_multigpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
timeout 300 $TR scripts/host_path_probe.py --mib 1024 --reps 2 > $O/r2g_host_path_probe.log 2> $O/r2g_host_path_probe.err
timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2g_scale_n8_peer.json 2> $O/r2g_scale_n8_peer.err
B200_HALO_NCCL=1 timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e > $O/r2g_scale_n8_nccl.json 2> $O/r2g_scale_n8_nccl.err
python - > $O/r2g_c5_8rank.log 2>&1 <<'PY'
import sys
sys.path.insert(0, "tests")
import compare_runs as cr
args = "--nx 8192 --ny 8192 --integrator dirk --order 3 --tf 1e-3 --nout 1 --output 1".split()
for np_ in (8, 1):
    wd, out = cr.run(cr.B200_BIN, args, np_, timeout=400)
    print("=== C5 8192^2 DIRK3 + PCG + Jacobi, %d rank(s)" % np_)
    print("\n".join(l for l in out.splitlines() if any(k in l for k in ("Total simulation", "Steps  ", "LS iters", "NLS iters  ", "B200", "Implicit RHS", "LS RHS"))))
PY
ls -la $O | tail -8
