#!/bin/bash
# round 2, call O (1 GPU): is the run-to-run spread of config c2 (2.2 .. 15 ms per step) lazy module loading?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
for i in 1 2 3 4; do CUDA_MODULE_LOADING=EAGER python bench.py --config c2 --no-cpu-baseline --no-e2e > $O/r2o_c2_eager_$i.json 2> $O/r2o_c2_eager_$i.err; done
for i in 1 2 3 4; do python bench.py --config c2 --no-cpu-baseline --no-e2e > $O/r2o_c2_lazy_$i.json 2> $O/r2o_c2_lazy_$i.err; done
for i in 1 2 3; do B200_NO_POLL=1 python bench.py --config c2 --no-cpu-baseline --no-e2e > $O/r2o_c2_nopoll_$i.json 2> $O/r2o_c2_nopoll_$i.err; done
ls $O | grep r2o | wc -l
