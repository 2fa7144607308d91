#!/bin/bash
# round 2, first GPU call (2 GPUs): the whole gpu test suite incl. the config-size and multi-GPU tests, bench at N=1 and N=2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err; echo "bench n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_c2_n2.json 2> gpurun_out/bench_c2_n2.err; echo "bench n2 rc=$?"
for f in gpurun_out/bench_c2_n1.json gpurun_out/bench_c2_n2.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    s=d.get("sharded") or {}
    print(sys.argv[1], "ms/step %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"], "e2e %.0f (%.2f of ceiling)"%(d["e2e"]["value"], d["e2e"]["frac_of_ceiling"]),
          "| sharded c4: %.3f ms, %.0f Mpix/s, exch %s us, e2e %s"%(s.get("ms_per_step",0), s.get("value",0), s.get("exchange_us"), (s.get("e2e") or {}).get("value")))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
tail -5 gpurun_out/bench_c2_n2.err
