#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/variants.txt
for v in default $VARIANTS; do
  if [ "$v" = default ]; then unset MORSI_CUDA_LIB; else export MORSI_CUDA_LIB=$PWD/build/variants/$v/libmorsi_cuda.so; fi
  timeout 300 python scratch/variant_check.py $v >> gpurun_out/variants.txt 2>&1 || echo "[$v] FAILED rc=$?" >> gpurun_out/variants.txt
done
unset MORSI_CUDA_LIB
cat gpurun_out/variants.txt
