#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/variants.txt
for v in default swap old; do
  if [ "$v" = default ]; then unset MORSI_CUDA_LIB; else export MORSI_CUDA_LIB=$PWD/build/variants/$v/libmorsi_cuda.so; fi
  timeout 120 python scratch/variant_check.py $v >> gpurun_out/variants.txt 2>&1 || echo "[$v] FAILED rc=$?" >> gpurun_out/variants.txt
done
unset MORSI_CUDA_LIB
cat gpurun_out/variants.txt
timeout 700 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py -m gpu -x -q --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
