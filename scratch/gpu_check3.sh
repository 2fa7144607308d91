#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
{
python scratch/pcie_bw.py
python scratch/time_op.py disk11 tophat 4096 4096 3 0
python scratch/time_op.py disk14 opening 4096 4096 3 0
python scratch/time_op.py cross opening 1920 1080 192 0
python scratch/time_op.py square tophat 1920 1080 192 0
python scratch/time_op.py square oscillation 1920 1080 192 0
python scratch/time_op.py cross gradient 1920 1080 192 0
for rows in 256 512 1024 2048; do echo "chunk rows $rows"; MORSI_CUDA_CHUNK_ROWS=$rows python bench.py --workload c2 --steps 5 --warmup 3 --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e', round(d['e2e']['value']))"; done
} 2>&1 | tee gpurun_out/timings3.txt
