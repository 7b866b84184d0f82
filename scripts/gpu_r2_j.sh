#!/bin/bash
# round 2, call J (8 GPUs): why was the N = 8 line of call G 65-67 ms per step against 54.6 at N = 1?
#   (1) the N = 8 line as the driver runs it (20 steps, 5 warm-up), with per-rank device time / SM clock / power and the
#       per-step diagnostic; (2) eight INDEPENDENT N = 1 benches at once (no exchange at all: what the box itself does
#       to eight busy GPUs); (3) one N = 1 bench alone on the same box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
B200_BENCH_PER_STEP=1 timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e > $O/r2j_scale_n8_peer.json 2> $O/r2j_scale_n8_peer.err
for i in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$i timeout 400 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/r2j_replica_$i.json 2> $O/r2j_replica_$i.err &
done
wait
timeout 400 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $O/r2j_n1_alone.json 2> $O/r2j_n1_alone.err
B200_HALO_NCCL=1 timeout 400 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e > $O/r2j_scale_n8_nccl.json 2> $O/r2j_scale_n8_nccl.err
ls -la $O | tail -14
# (same box, one GPU) small grids with the adaptive rows of the one-stage kernels, host profile
D=./ceda-demonstrations_b200/bin/diffusion_2D_b200
{
for n in 32 128; do
echo "=== ${n}^2 rkc tf=1"
for rep in 1 2 3 4; do timeout 300 $D --nx $n --ny $n --integrator rkc --tf 1 --nout 1 --output 1 | grep -E "Total simulation|B200 kernel launches"; done
B200_HOST_PROFILE=1 timeout 300 $D --nx $n --ny $n --integrator rkc --tf 1 --nout 1 --output 1 2>&1 | grep -E "Total simulation|host profile"
done
} > $O/r2j_small_grids.log 2>&1
