#!/bin/bash
mkdir -p gpurun_out
cap() {  # name regex skip command...
  local name=$1 re=$2 skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$re" -s $skip -c 1 -o /tmp/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
  python scratch/ncu_summary.py /tmp/$name.ncu-rep > gpurun_out/${name}_summary.txt 2>&1
  head -4 gpurun_out/${name}_summary.txt
}
cap kdisk_c2 '^k_disk$' 4 python scratch/time_op.py disk7 opening 4096 4096 3 0 3
cap kdisk_c4 '^k_disk$' 3 python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu
cap kdisk_erosion '^k_disk$' 4 python scratch/time_op.py disk7 erosion 4096 4096 3 0 3
python scratch/make_traffic.py c2=kdisk_c2 c4=kdisk_c4 c3=kmedian_c3 c5=ksmall_c5
