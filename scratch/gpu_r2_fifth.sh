#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py::test_runtime_rowrun_shapes tests/test_gpu_pipeline.py::test_cli_all_and_streaming tests/test_gpu_fuzz.py -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
for args in "disk20 closing 4096 4096 3 0" "disk20 erosion 4096 4096 3 0" "disk16 tophat 4096 4096 3 0" "disk6.5 opening 4096 4096 3 0" "hrec31 dilation 4096 4096 3 0" "disk32 gradient 4096 4096 1 0"; do
  timeout 120 python scratch/time_op.py $args 10 2>&1 | tail -1
  MORSI_RUNS=0 timeout 120 python scratch/time_op.py $args 5 2>&1 | tail -1 | sed 's/^/   (k_tiled) /'
done | tee gpurun_out/runs_timings.txt
