#!/bin/bash
mkdir -p gpurun_out
name=kdual_disk7
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:^k_disk_dual$" -s 4 -c 1 -o /tmp/$name -f python scratch/time_op.py disk7 gradient 4096 4096 3 0 3 > gpurun_out/ncu_$name.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
python scratch/ncu_summary.py /tmp/$name.ncu-rep > gpurun_out/${name}_summary.txt 2>&1
cat gpurun_out/${name}_summary.txt
