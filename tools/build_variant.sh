#!/bin/bash
# developer helper: build a library variant with extra nvcc defines: tools/build_variant.sh NAME "-DMX_K3_MINBLOCKS=4"
# -> build/variants/libmarxb200_NAME.so, used through MARXB200_LIB=... (marx_b200/api.py).  ONLY=kernels recompiles kernels.cu alone
# and links the other objects of the regular build.
set -e
cd "$(dirname "$0")/../marx_b200/csrc"
NAME=$1; shift
OUT=../../build/variants; mkdir -p $OUT
FMAD=${FMAD:-false}
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=$FMAD -std=c++17 -Xcompiler -fPIC $@"
UNITS="kernels level1_kernels aspsol_kernels pileup_kernels marxb200 comm"
[ -n "$ONLY" ] && UNITS="$ONLY"
OBJS=""
for f in kernels level1_kernels aspsol_kernels pileup_kernels marxb200 comm; do
  if echo " $UNITS " | grep -q " $f "; then nvcc $FL -c $f.cu -o $OUT/${f}_$NAME.o & OBJS="$OBJS $OUT/${f}_$NAME.o"; else OBJS="$OBJS ../../build/$f.o"; fi
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libmarxb200_$NAME.so $OBJS ../../build/calpack.o ../../build/writer.o -ldl -lpthread
echo built $OUT/libmarxb200_$NAME.so
