#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fuzz.py "tests/test_gpu_parity.py::test_disk_kernels_multi_strip_multi_band" tests/test_gpu_parity.py::test_disk_kernels_unaligned_width "tests/test_gpu_parity.py::test_adversarial_golden" -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
for op in gradient laplacian cblur enhance; do
  timeout 100 python scratch/time_op.py disk7 $op 4096 4096 3 0 20 | tail -1
  MORSI_DISK_DUAL=0 timeout 100 python scratch/time_op.py disk7 $op 4096 4096 3 0 20 | tail -1 | sed 's/^/   (two-role kernel) /'
done | tee gpurun_out/dual_timings.txt
for e in disk5 disk3 disk2.5; do timeout 100 python scratch/time_op.py $e gradient 4096 4096 3 0 20 | tail -1; MORSI_DISK_DUAL=0 timeout 100 python scratch/time_op.py $e gradient 4096 4096 3 0 20 | tail -1 | sed 's/^/   (two-role kernel) /'; done | tee -a gpurun_out/dual_timings.txt
