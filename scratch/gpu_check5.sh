#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
{
python scratch/time_op.py disk5 rank 8192 8192 1 0 3
MORSI_TILED=0 python scratch/time_op.py disk5 rank 8192 8192 1 0 2
python scratch/time_op.py dysk7 rank 4096 4096 3 0 3
} 2>&1 | tee gpurun_out/timings5.txt
