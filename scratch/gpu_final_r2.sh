#!/bin/bash
# end-of-round-2 evidence on ONE B200: the whole gpu test suite, ncu launch list + full captures (CSV exports),
# bench lines of all five workloads, the reference arm, smoke, the signed-zero cliff
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sharded > gpurun_out/ncu_launch.log 2>&1
cap() {  # name regex skip command...
  local name=$1 re=$2 skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$re" -s $skip -c 1 -o /tmp/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
  python scratch/ncu_summary.py /tmp/$name.ncu-rep > gpurun_out/${name}_summary.txt 2>&1
  head -3 gpurun_out/${name}_summary.txt
}
cap kdisk_c2 '^k_disk$' 4 python scratch/time_op.py disk7 opening 4096 4096 3 0 3
cap kdisk_c4 '^k_disk$' 3 python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu
cap kmedian_c3 k_median_quad 3 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu --no-sharded
cap kdual_disk7 '^k_disk_dual$' 4 python scratch/time_op.py disk7 gradient 4096 4096 3 0 3
cap kruns_disk20 k_runs 4 python scratch/time_op.py disk20 erosion 4096 4096 3 0 3
cap ksmall_c5 k_small 3 python bench.py --workload c5 --steps 1 --warmup 3 --no-cpu --no-sharded
python scratch/make_traffic.py c2=kdisk_c2 c4=kdisk_c4 c3=kmedian_c3 c5=ksmall_c5
cp gpurun_out/traffic.json profiles/traffic.json
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
for w in c1 c3 c5; do timeout 400 python bench.py --workload $w --steps 20 --warmup 3 --no-sharded > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; done
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
for w in c2 c1 c3 c4 c5; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_$w.json") if l.startswith("{")][-1])
    print("$w", "ms/step %.4f"%d["ms_per_step"], "Mpix/s %.0f"%d["value"], "frac %.3f"%d["roofline"]["frac"], "alu", (d["roofline"].get("alu_pipe") or {}).get("frac"), "e2e", d["e2e"] and round(d["e2e"]["value"]), "cli", (d["e2e"] or {}).get("cli"), "launches", d["gpu_launches"])
except Exception as ex: print("$w unreadable", ex)
PY
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cut -c1-200 gpurun_out/bench_ref.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
# the signed-zero cliff: the same operations on data with 1 % -0.0 (dist 2) against clean data (dist 0)
for args in "disk7 opening 4096 4096 3" "disk7 erosion 4096 4096 3" "cross gradient 1920 1080 192" "disk20 erosion 4096 4096 1"; do
  timeout 120 python scratch/time_op.py $args 0 5 | tail -1; timeout 300 python scratch/time_op.py $args 2 3 | tail -1
  MORSI_TILED_EXACT=0 timeout 300 python scratch/time_op.py $args 2 2 | tail -1 | sed 's/^/   (k_exact re-run) /'
done | tee gpurun_out/negzero_cliff.txt
du -sh gpurun_out
