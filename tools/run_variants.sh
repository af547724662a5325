#!/bin/bash
# developer script (GPU box): bench every build/variants/libmphx_*.so, print the kernel-group split
for so in build/variants/libmphx_*.so; do
  MPHX_LIB=$PWD/$so python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 --state lattice ${BENCH_ARGS:-} 2>&1 | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1]); p=d['roofline']['kernel_ms_per_step']
print('$so', round(d['ms_per_step'],3), {k: round(v,3) for k,v in p.items()})"
done
