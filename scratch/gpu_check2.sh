#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
{
python scratch/time_op.py disk7 opening 4093 4096 3 0
MORSI_DISK=0 MORSI_TILED=0 python scratch/time_op.py disk7 opening 4093 4096 3 0 3
python scratch/time_op.py disk7 opening 4096 4096 3 0
python scratch/time_op.py disk7 oscillation 4093 4096 3 0
python scratch/time_op.py dysk7 erosion 4096 4096 3 0
MORSI_TILED=0 python scratch/time_op.py dysk7 erosion 4096 4096 3 0 3
python scratch/time_op.py hrec40 dilation 4096 4096 3 0
MORSI_TILED=0 python scratch/time_op.py hrec40 dilation 4096 4096 3 0 3
python scratch/time_op.py disk11 tophat 4096 4096 3 0 3
MORSI_TILED=0 python scratch/time_op.py disk11 tophat 4096 4096 3 0 2
python scratch/time_op.py drec9 gradient 4096 4096 3 0
python scratch/time_op.py disk20 closing 4096 4096 1 0 2
} 2>&1 | tee gpurun_out/timings_generic.txt
for w in c2 c4; do python bench.py --workload $w --steps 20 --warmup 3 --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w ms/step %.4f frac %.3f'%(d['ms_per_step'],d['roofline']['frac']), d['roofline'].get('alu_pipe'))"; done
