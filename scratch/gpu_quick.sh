#!/bin/bash
# quick GPU check: selected parity tests + bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for w in ${WORKLOADS:-c2 c1 c3 c4 c5}; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$w.json"))
print("$w", "ms/step %.4f"%d["ms_per_step"], "Mpix/s %.0f"%d["value"], "frac %.3f"%d["roofline"]["frac"], "e2e", d["e2e"] and round(d["e2e"]["value"]), "launches", d["gpu_launches"])
PY
done
