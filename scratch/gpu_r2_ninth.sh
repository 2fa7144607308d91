#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fuzz.py "tests/test_gpu_parity.py::test_disk_kernels_multi_strip_multi_band" tests/test_gpu_parity.py::test_disk_kernels_unaligned_width "tests/test_gpu_parity.py::test_adversarial_golden" tests/test_gpu_config_sizes.py::test_c2_crops_all_planes tests/test_gpu_parity.py::test_full_size_properties tests/test_gpu_parity.py::test_device_and_band_entry_points -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for op in opening closing tophat bothat oscillation; do
  timeout 100 python scratch/time_op.py disk7 $op 4096 4096 3 0 20 | tail -1
  MORSI_DISK_TWO1=0 timeout 100 python scratch/time_op.py disk7 $op 4096 4096 3 0 20 | tail -1 | sed 's/^/   (two-role kernel) /'
done | tee gpurun_out/two1_timings.txt
for e in disk5 disk3; do timeout 100 python scratch/time_op.py $e opening 4096 4096 3 0 20 | tail -1; MORSI_DISK_TWO1=0 timeout 100 python scratch/time_op.py $e opening 4096 4096 3 0 20 | tail -1 | sed 's/^/   (two-role kernel) /'; done | tee -a gpurun_out/two1_timings.txt
