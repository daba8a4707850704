#!/bin/bash
# Step-level A/B of library variants on ONE box: tools/ab_bench.sh v1 v2 ...  (variant = name under tools/ubench/_bin, or "." for
# the in-tree library), two interleaved rounds of `bench.py --steps 10 --warmup 3`; prints seq/s, ms/step, e2e, GEMM frac, SM MHz.
cd "$(dirname "$0")/.."
for r in 1 2; do for v in "$@"; do
  if [ "$v" = "." ]; then L=$PWD/transformerscandobayesianinference_b200/libpfn_b200.so; else L=$PWD/tools/ubench/_bin/libpfn_$v.so; fi
  PFN_B200_LIB=$L timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"
done; done
