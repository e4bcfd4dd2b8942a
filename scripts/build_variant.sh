#!/bin/bash
# Build a variant of the library with extra -D flags for ONE translation unit: build/lib_<tag>.so
# usage: bash scripts/build_variant.sh <tag> <tu (ffb_rd|ffb_kernels|ffb_staged|ffb_wide|ffb_wide2|ffb_train)> [-DFLAG ...]
set -e
tag=$1; tu=$2; shift 2
mkdir -p build/obj
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Iinclude -Iflowfusion_b200/csrc -Xcompiler -fPIC"
nvcc $FL "$@" -c flowfusion_b200/csrc/$tu.cu -o build/obj/${tu}_$tag.o
objs=""
for t in ffb_kernels ffb_staged ffb_rd ffb_wide ffb_wide2 ffb_train; do
  if [ $t == $tu ]; then objs="$objs build/obj/${tu}_$tag.o"; else objs="$objs build/obj/$t.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/lib_$tag.so $objs
echo built build/lib_$tag.so
