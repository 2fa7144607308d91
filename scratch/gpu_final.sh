#!/bin/bash
# end-of-round evidence: parity tests, ncu launch list + full captures (CSV exports), bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
cap() { bash scratch/gpu_ncu_one.sh $1 $2 $3 $4 > /dev/null 2>&1; }
cap kdisk_c2 k_disk 6 c2
cap kdisk_c4 k_disk 3 c4
cap kmedian_c3 k_median_quad 3 c3
cap ksmall_c5 k_small 3 c5
cap ksmall_c1 k_small 3 c1
python scratch/make_traffic.py c2=kdisk_c2 c4=kdisk_c4 c3=kmedian_c3 c5=ksmall_c5 c1=ksmall_c1
mkdir -p profiles; cp gpurun_out/traffic.json profiles/traffic.json
for w in c2 c1 c3 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$w.json"))
print("$w", "ms/step %.4f"%d["ms_per_step"], "Mpix/s %.0f"%d["value"], "frac %.3f"%d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "e2e", d["e2e"] and round(d["e2e"]["value"]), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],3), "launches", d["gpu_launches"])
PY
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()"
rm -f gpurun_out/*_source.csv.bak; du -sh gpurun_out
# extra: the fused 3x3 two-stage kernel (cross opening of the C5 batch)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_small_2 -s 3 -c 1 -o /tmp/ksmall2 -f python scratch/time_op.py cross opening 1920 1080 192 0 3 > gpurun_out/ncu_ksmall2.log 2>&1
ncu -i /tmp/ksmall2.ncu-rep --page source --csv > gpurun_out/ksmall2_source.csv 2>/dev/null
python scratch/ncu_summary.py /tmp/ksmall2.ncu-rep > gpurun_out/ksmall2_summary.txt 2>&1
for op in erosion opening gradient; do MORSI_CUDA_PATH=fast python scratch/time_op.py disk7 $op 4096 4096 3 0 20; done
