#!/bin/bash
# usage: tools/build_fused_variant.sh <tag> <f64|f32> <group 0-3> "<extra nvcc flags>"  -> mjhmc_b200/_variants/lib_<tag>.so
# recompiles ONE register-kernel instantiation unit with extra flags and links it with the objects of the last full build
tag=$1; tn=$2; g=$3; flags=$4
das=(1 3 6 10); dbs=(2 4 8 16)
ct=double; [ "$tn" = f32 ] && ct=float
mkdir -p mjhmc_b200/_variants
unit=fused_inst_${tn}_g$g
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $flags \
  -DMJ_T=$ct -DMJ_TAG=${tn}_g$g -DMJ_DA=${das[$g]} -DMJ_DB=${dbs[$g]} \
  -c mjhmc_b200/csrc/fused_inst.cu -o mjhmc_b200/_variants/${unit}_$tag.o || exit 1
objs=$(ls mjhmc_b200/_build/*.o | grep -v "/$unit.o")
/usr/local/cuda/bin/nvcc -shared -o mjhmc_b200/_variants/lib_$tag.so $objs mjhmc_b200/_variants/${unit}_$tag.o -gencode arch=compute_100a,code=sm_100a
