#!/bin/bash
mkdir -p gpurun_out
run() { wl=$1; shift; echo "== $wl $*"; env "$@" python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms/step %.4f frac %.3f launches %d'%(d['ms_per_step'],d['roofline']['frac'],d['gpu_launches']))"; }
{
for b in 8 9 12 16 17 18 26; do
run c2 MORSI_DISK_W=2 MORSI_DISK_BANDS=$b
done
for b in 2 3 4 5 6; do
run c2 MORSI_DISK_W=4 MORSI_DISK_BANDS=$b
done
for b in 12 14 16 18 25 32; do
run c4 MORSI_DISK_W=4 MORSI_DISK_BANDS=$b
done
} 2>&1 | tee gpurun_out/sweep_disk_bands.txt
