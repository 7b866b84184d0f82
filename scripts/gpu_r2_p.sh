#!/bin/bash
# round 2, call P (1 GPU): the library built from several translation units (chain kernels one depth per object);
# SPLIT flavour of k_chain_march: bit-identity on hardware, A/B on the headline (alternating, same box), ncu of one launch
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -m gpu -q -x -k "split or head or temporal or variants or stage_chain or golden" 2>&1 | tail -6 > $O/r2p_pytest.log
for i in 1 2 3; do
  B200_CHAIN_SPLIT=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2p_c3_plain_$i.json 2> $O/r2p_c3_plain_$i.err
  B200_CHAIN_SPLIT=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2p_c3_split_$i.json 2> $O/r2p_c3_split_$i.err
done
B200_CHAIN_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_march --launch-skip 45 --launch-count 2 \
  -o $O/r2p_chain4_split_head_and_body -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2p_ncu_full.log 2>&1
ls -la $O | tail -4
