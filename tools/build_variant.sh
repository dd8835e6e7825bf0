#!/bin/bash
# Build a variant of the product library next to the default one (A/B experiments on the GPU box):
#   tools/build_variant.sh <suffix> <extra nvcc flags...>    -> ecrad_b200/libecrad_b200_<suffix>.so
set -e
SUF=$1; shift
SRC=$(cd "$(dirname "$0")/../ecrad_b200/csrc" && pwd)
OBJ=/tmp/ecb_variant_$SUF
mkdir -p $OBJ
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in api kernels ecckd solver_sw solver_lw solver_tc solver_sp solver_scan gas_band; do
  ( /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-O2 -Xptxas -v "$@" -c $SRC/$f.cu -o $OBJ/$f.o 2> $OBJ/$f.log || (cat $OBJ/$f.log; exit 1) ) &
done
wait
/usr/local/cuda/bin/nvcc $ARCH -shared -o $SRC/../libecrad_b200_$SUF.so $OBJ/*.o -lcudart
ls -la $SRC/../libecrad_b200_$SUF.so
