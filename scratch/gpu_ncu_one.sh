#!/bin/bash
# usage: gpu_ncu_one.sh name kernel-regex skip workload   -> summary + source csv in gpurun_out/
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/$1 -f python bench.py --workload $4 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_$1.log 2>&1
ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
python scratch/ncu_summary.py /tmp/$1.ncu-rep > gpurun_out/$1_summary.txt 2>&1
cat gpurun_out/$1_summary.txt
