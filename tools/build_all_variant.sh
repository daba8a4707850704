#!/bin/bash
# Whole-library A/B build with extra -D macros on EVERY source: tools/build_all_variant.sh <name> "<flags>" -> tools/ubench/_bin/libpfn_<name>.so
set -e
cd "$(dirname "$0")/.."
CS=transformerscandobayesianinference_b200/csrc
OUT=tools/ubench/_bin
mkdir -p $OUT/obj_$1
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $2"
OBJS=""
for f in runtime optimizer gemm_tc gemm_tc_c2g gemm_simt rowwise bar_nll attention_simt attention_tc attention_bwd_tc attention_bwd_dq gp_sampler dropout; do
  nvcc $FLAGS -c $CS/$f.cu -o $OUT/obj_$1/$f.o &
  OBJS="$OBJS $OUT/obj_$1/$f.o"
done
wait
nvcc -shared -o $OUT/libpfn_$1.so $OBJS -gencode arch=compute_100a,code=sm_100a
ls -la $OUT/libpfn_$1.so
