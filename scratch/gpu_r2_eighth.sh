#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fuzz.py "tests/test_gpu_parity.py::test_disk_kernels_multi_strip_multi_band" tests/test_gpu_parity.py::test_disk_kernels_unaligned_width "tests/test_gpu_parity.py::test_adversarial_golden" tests/test_gpu_parity.py::test_vs_oracle_shapes_elements_ops -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for op in gradient laplacian cblur igradient eblur erosion opening; do
  timeout 100 python scratch/time_op.py disk7 $op 4096 4096 3 0 20 | tail -1
done | tee gpurun_out/dual_timings2.txt
timeout 100 python scratch/time_op.py disk15 igradient 40000 10000 1 0 5 | tail -1 | tee -a gpurun_out/dual_timings2.txt
