#!/bin/bash
# ncu captures, exported to CSV on the box (the .ncu-rep files are too large to bring back)
mkdir -p gpurun_out
./scratch/ubench_alu > gpurun_out/ubench_alu.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
cap() { # name kernel-regex skip workload
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o /tmp/$1 -f python bench.py --workload $4 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  python scratch/ncu_summary.py /tmp/$1.ncu-rep > gpurun_out/$1_summary.txt 2>&1
}
cap kdisk_c2 k_disk 6 c2
cap kdisk_c4 k_disk 3 c4
cap kmedian_c3 k_median 3 c3
cap ksmall_c5 k_small 3 c5
cap ksmall_c1 k_small 3 c1
ls -la gpurun_out; du -sh gpurun_out
