#!/bin/bash
# last call of the round: full parity suite, bench lines, k_small captures (the other ncu summaries are current)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
bash scratch/gpu_ncu_one.sh ksmall_c5 k_small 3 c5 > /dev/null 2>&1
bash scratch/gpu_ncu_one.sh ksmall_c1 k_small 3 c1 > /dev/null 2>&1
for w in c2 c1 c3 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$w.json"))
print("$w", "ms/step %.4f"%d["ms_per_step"], "Mpix/s %.0f"%d["value"], "frac %.3f"%d["roofline"]["frac"], "e2e", d["e2e"] and round(d["e2e"]["value"]), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],3), "launches", d["gpu_launches"])
PY
done
python -c "import __graft_entry__ as g; g.smoke()"
