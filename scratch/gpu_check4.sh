#!/bin/bash
mkdir -p gpurun_out
{
for rpw in 64 32 16 8; do
echo "rpw $rpw"; MORSI_SMALL_RPW=$rpw python scratch/time_op.py cross opening 1920 1080 192 0
MORSI_SMALL_RPW=$rpw python scratch/time_op.py square tophat 1920 1080 192 0
done
python scratch/time_op.py square oscillation 1920 1080 192 0
python scratch/time_op.py cross gradient 1920 1080 192 0
python scratch/time_op.py square closing 1024 1024 1 0 50
} 2>&1 | tee gpurun_out/timings4.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "vs_oracle or adversarial or known or nan or multi_plane" 2>&1 | tail -3
