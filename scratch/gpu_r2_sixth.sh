#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py::test_runtime_rowrun_shapes -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for args in "disk20 closing 4096 4096 3 0" "disk20 closing 4096 4096 1 0" "disk20 erosion 4096 4096 3 0" "disk16 tophat 4096 4096 3 0" "disk6.5 opening 4096 4096 3 0" "disk32 gradient 4096 4096 1 0"; do
  timeout 120 python scratch/time_op.py $args 10 2>&1 | tail -1
done | tee gpurun_out/runs_timings.txt
