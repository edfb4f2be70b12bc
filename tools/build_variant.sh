#!/bin/bash
# developer helper: build a library variant with extra nvcc defines: tools/build_variant.sh NAME "-DMX_STAGE_TILE=64"
set -e
cd "$(dirname "$0")/../marx_b200/csrc"
NAME=$1; shift
OUT=../../build/variants; mkdir -p $OUT
FMAD=${FMAD:-false}
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=$FMAD -std=c++17 -Xcompiler -fPIC $@"
nvcc $FL -c kernels.cu -o $OUT/kernels_$NAME.o
nvcc $FL -c level1_kernels.cu -o $OUT/level1_kernels_$NAME.o
nvcc $FL -c aspsol_kernels.cu -o $OUT/aspsol_kernels_$NAME.o
nvcc $FL -c marxb200.cu -o $OUT/marxb200_$NAME.o
g++ -O2 -std=c++17 -fPIC -c calpack.cpp -o $OUT/calpack_$NAME.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libmarxb200_$NAME.so $OUT/kernels_$NAME.o $OUT/level1_kernels_$NAME.o $OUT/aspsol_kernels_$NAME.o $OUT/marxb200_$NAME.o $OUT/calpack_$NAME.o
echo built $OUT/libmarxb200_$NAME.so
