#!/bin/bash
mkdir -p gpurun_out
run() { wl=$1; shift; echo "== $wl $*"; env "$@" python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms/step %.4f frac %.3f launches %d clocks %s'%(d['ms_per_step'],d['roofline']['frac'],d['gpu_launches'],d['clocks']))"; }
{
run c2 X=1
run c2 MORSI_DISK_W=2
for b in 3 4 5 6 8 9 10 12 13; do
run c2 MORSI_DISK_W=4 MORSI_DISK_BANDS=$b
done
run c4 X=1
} 2>&1 | tee gpurun_out/sweep_disk_ng.txt
python -m pytest tests -m gpu -x -q -k "disk_kernels" 2>&1 | tail -2
