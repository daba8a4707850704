#!/bin/bash
# A/B of one environment knob on ONE box: tools/ab_env.sh "<command>" VAR=a VAR=b ...  (each run 2x, interleaved)
cd "$(dirname "$0")/.."
CMD="$1"; shift
for rep in 1 2; do for kv in "$@"; do echo "== $kv (rep $rep)"; env $kv bash -c "$CMD"; done; done
