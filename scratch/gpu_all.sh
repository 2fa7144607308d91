#!/bin/bash
# one GPU-box call: parity tests, bench lines for all workloads, launch list, ncu full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for w in c2 c1 c3 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  cat gpurun_out/bench_$w.json
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_disk -s 6 -c 1 -o gpurun_out/kdisk_c2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_disk -s 3 -c 1 -o gpurun_out/kdisk_c4 -f python bench.py --workload c4 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_c4.log 2>&1
ls -la gpurun_out
