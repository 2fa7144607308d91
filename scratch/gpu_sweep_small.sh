#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c5 ms/step %.4f frac %.3f'%(d['ms_per_step'],d['roofline']['frac']))"; }
{
for rpw in 8 16 32 64; do
run MORSI_SMALL_PF=3 MORSI_SMALL_WX=3 MORSI_SMALL_RPW=$rpw
done
for rpw in 8 16 32; do
run MORSI_SMALL_PF=3 MORSI_SMALL_WX=0 MORSI_SMALL_RPW=$rpw
done
run MORSI_SMALL_PF=3 MORSI_SMALL_WX=1 MORSI_SMALL_RPW=32
run MORSI_SMALL_PF=6 MORSI_SMALL_WX=3 MORSI_SMALL_RPW=16
} 2>&1 | tee gpurun_out/sweep_small.txt
