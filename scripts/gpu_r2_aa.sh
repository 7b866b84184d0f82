#!/bin/bash
# round 2, call AA (1 GPU): which flavour should be the default?  On ONE box, alternating, three repetitions:
# (make ab builds build/ab/libb200sts_noxshare.so)
# {plain cp.async ring, BULK ring with 4 rows in flight} x {x-direction products shared, not shared (A/B build of the
# library, -DB200_NO_XSHARE)}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
L=ceda-demonstrations_b200/lib/libb200sts.so
cp $L build/ab/libb200sts_xshare.so
for i in 1 2 3; do
  for lib in xshare noxshare; do
    cp build/ab/libb200sts_$lib.so $L
    for bulk in 0 2; do
      B200_CHAIN_BULK=$bulk python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/r2aa_${lib}_bulk${bulk}_$i.json 2> $O/r2aa_${lib}_bulk${bulk}_$i.err
    done
  done
done
cp build/ab/libb200sts_xshare.so $L
ls $O | grep -c r2aa_
