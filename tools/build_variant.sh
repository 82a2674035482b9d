#!/bin/bash
# usage: tools/build_variant.sh <tag> <unit> "<extra nvcc flags>"   -> mjhmc_b200/_variants/lib_<tag>.so
# recompiles ONE translation unit with extra flags and links it with the objects of the last full build
tag=$1; unit=$2; flags=$3
mkdir -p mjhmc_b200/_variants
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $flags \
  -c mjhmc_b200/csrc/$unit.cu -o mjhmc_b200/_variants/${unit}_$tag.o || exit 1
objs=$(ls mjhmc_b200/_build/*.o | grep -v "/$unit.o")
/usr/local/cuda/bin/nvcc -shared -o mjhmc_b200/_variants/lib_$tag.so $objs mjhmc_b200/_variants/${unit}_$tag.o -gencode arch=compute_100a,code=sm_100a
