#!/bin/bash
# N-GPU box: the multi-GPU parity tests + bench at N (C2 weak + the C4 sharded sub-record)
N=${N:-4}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 240 > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -8 gpurun_out/pytest_multi.log
grep -c "bit-identical" gpurun_out/mgpu_check.log; grep -i "mismatch" gpurun_out/mgpu_check.log | head -5
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_c2_n$N.json 2> gpurun_out/bench_c2_n$N.err; echo "bench rc=$?"
python - gpurun_out/bench_c2_n$N.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    s=d.get("sharded") or {}
    print("C2 ms/step %.4f e2e %.0f (%.2f of ceiling %.0f)"%(d["ms_per_step"], d["e2e"]["value"], d["e2e"]["frac_of_ceiling"], d["e2e"]["ceiling"]["value"]))
    print("sharded c4 n=%d: %.3f ms, %.0f Mpix/s, exch %s us, e2e %s of ceiling %s"%(s.get("n_gpus",0), s.get("ms_per_step",0), s.get("value",0), s.get("exchange_us"), (s.get("e2e") or {}).get("value"), (s.get("e2e") or {}).get("ceiling")))
except Exception as ex:
    print("unreadable:", ex); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
