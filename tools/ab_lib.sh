#!/bin/bash
# A/B of library build variants: tools/ab_lib.sh "<python command>" v1 v2 ...   (variants built by tools/build_variants.sh)
cd "$(dirname "$0")/.."
CMD="$1"; shift
for v in "$@"; do
  echo "== $v"; PFN_B200_LIB=$PWD/tools/ubench/_bin/libpfn_$v.so timeout 300 $CMD
done
