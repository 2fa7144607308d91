#!/bin/bash
# usage: build_variant.sh NAME "[extra nvcc -D flags]" [k_disk source]
# builds build/variants/NAME/libmorsi_cuda.so with only disk7 / disk15 (shapes 8, 13) in k_disk, reusing the other objects
set -e
name=$1; flags=$2; src=${3:-imscript_b200/csrc/k_disk.cu}
out=build/variants/$name; mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iimscript_b200/csrc -DMORSI_DISK_IDS="T(8) T(13)" $flags -c $src -o $out/k_disk.o
objs=$(ls build/obj/*.o | grep -v "k_disk\|iio\|k_march\|k_stubs")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libmorsi_cuda.so $out/k_disk.o $objs -lm
echo "built $out/libmorsi_cuda.so"
