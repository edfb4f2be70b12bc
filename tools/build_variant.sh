#!/bin/bash
# developer helper: build a library variant with extra nvcc defines: tools/build_variant.sh NAME "-DMX_MATH=0"
set -e
cd "$(dirname "$0")/../marx_b200/csrc"
NAME=$1; shift
OUT=../../build/variants; mkdir -p $OUT
FMAD=${FMAD:-false}
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=$FMAD -std=c++17 -Xcompiler -fPIC $@"
for f in kernels level1_kernels aspsol_kernels pileup_kernels marxb200 comm; do
  nvcc $FL -c $f.cu -o $OUT/${f}_$NAME.o &
done
g++ -O2 -std=c++17 -fPIC -c calpack.cpp -o $OUT/calpack_$NAME.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libmarxb200_$NAME.so $OUT/kernels_$NAME.o $OUT/level1_kernels_$NAME.o \
  $OUT/aspsol_kernels_$NAME.o $OUT/pileup_kernels_$NAME.o $OUT/marxb200_$NAME.o $OUT/comm_$NAME.o $OUT/calpack_$NAME.o -ldl
echo built $OUT/libmarxb200_$NAME.so
