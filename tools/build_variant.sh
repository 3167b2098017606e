#!/bin/bash
# Builds cmda_b200/variants/lib_<name>.so: the library with voxel_factored.cu compiled under extra -D flags
# (kernel-shape sweeps without touching the shipped build).  Use: CMDA_B200_LIB=cmda_b200/variants/lib_<name>.so python bench.py ...
#   tools/build_variant.sh rows40 "-DCMDA_BAND_ROWS=40"      (CMDA_BAND_V2_UNROLL, CMDA_BAND_UNROLL, ... : voxel_factored.cu)
set -e
cd "$(dirname "$0")/../cmda_b200/csrc"
name=$1; shift
mkdir -p build/variants ../variants
make -j8 > /dev/null
unit=${UNIT:-voxel_factored}      # UNIT=pseudo_events tools/build_variant.sh ... rebuilds that translation unit instead
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -cudart static \
     -Xptxas -v "$@" -c -o build/variants/${unit}_$name.o $unit.cu 2> build/variants/$name.ptxas.log
objs=$(ls build/*.o | grep -v $unit.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xcompiler -fPIC -o ../variants/lib_$name.so $objs build/variants/${unit}_$name.o
echo built ../variants/lib_$name.so
