#!/bin/bash
# A/B builds of libpfn_b200.so with extra -D macros on ONE source file (select at run time with PFN_B200_LIB=<.so>).
#   usage: tools/build_variants.sh <file.cu> name1 "-DFLAG_A" name2 "-DFLAG_B" ...   ->  tools/ubench/_bin/libpfn_<name>.so
set -e
cd "$(dirname "$0")/.."
CS=transformerscandobayesianinference_b200/csrc
OUT=tools/ubench/_bin
mkdir -p $OUT/obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
VAR=$(basename $1 .cu); shift
ALL="runtime gemm_tc gemm_tc_c2g gemm_simt rowwise bar_nll attention_simt attention_tc attention_bwd_tc attention_bwd_dq gp_sampler dropout optimizer"
for f in $ALL; do
  [ $f = $VAR ] && continue
  [ $OUT/obj/$f.o -nt $CS/$f.cu ] || nvcc $FLAGS -c $CS/$f.cu -o $OUT/obj/$f.o &
done
wait
while [ $# -ge 2 ]; do
  ( nvcc $FLAGS $2 -c $CS/$VAR.cu -o $OUT/obj/${VAR}_$1.o
    OBJS=""; for f in $ALL; do if [ $f = $VAR ]; then OBJS="$OBJS $OUT/obj/${VAR}_$1.o"; else OBJS="$OBJS $OUT/obj/$f.o"; fi; done
    nvcc -shared -o $OUT/libpfn_$1.so $OBJS -gencode arch=compute_100a,code=sm_100a ) &
  shift 2
done
wait
ls -la $OUT/*.so
