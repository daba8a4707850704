#!/bin/bash
# A/B of dQ-kernel build variants on one GPU: tools/ab_dq.sh v1 v2 ...  (variants: tools/build_variants.sh attention_bwd_dq.cu ...)
cd "$(dirname "$0")/.."
for v in "$@"; do
  echo "== $v"; PFN_B200_LIB=$PWD/tools/ubench/_bin/libpfn_$v.so timeout 120 python tools/time_kernels.py attnparts | grep "part dq"
done
