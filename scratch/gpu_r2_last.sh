#!/bin/bash
# final verification of the committed state: the whole gpu suite + the default bench line + smoke
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_c2.json") if l.startswith("{")][-1])
print("c2 ms/step %.4f frac %.3f e2e %.0f cli %s sharded %.3f ms launches %d"%(d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], (d["e2e"].get("cli") or {}).get("value"), d["sharded"]["ms_per_step"], d["gpu_launches"]))
PY
timeout 120 python -c "import __graft_entry__ as g; g.smoke()"
