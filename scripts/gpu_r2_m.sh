#!/bin/bash
# round 2, call M (1 GPU): config c2 went from 2.30 to 9.34 ms per step between calls I and L -- which switch?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
for v in default B200_NO_POLL B200_NO_CHAIN_HEAD B200_NO_SPEC_EWT; do
  if [ $v = default ]; then E=""; else E="$v=1"; fi
  env $E B200_HOST_PROFILE=1 python bench.py --config c2 --no-cpu-baseline --no-e2e > $O/r2m_c2_$v.json 2> $O/r2m_c2_$v.err
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2m_c3_default.json 2> $O/r2m_c3_default.err
B200_NO_POLL=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2m_c3_nopoll.json 2> $O/r2m_c3_nopoll.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/r2m_c3_default_again.json 2> $O/r2m_c3_default_again.err
ls -la $O | tail -5
