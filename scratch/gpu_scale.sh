#!/bin/bash
# run on an N-GPU box: C4 strong scaling and C5 weak scaling
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/scale_gpus.txt
run() { n=$1; wl=$2; port=$((29500 + n)); 
  if [ $n -eq 1 ]; then timeout 300 python bench.py --gpus 1 --workload $wl --steps 10 --warmup 3 --no-cpu
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --workload $wl --steps 10 --warmup 3 --no-cpu 2> gpurun_out/scale_${wl}_$n.err | grep '^{'
  fi; }
for n in ${NS:-8 4 2 1}; do run $n c4 | tee gpurun_out/scale_c4_$n.json; done
for n in ${NS5:-8 1}; do run $n c5 | tee gpurun_out/scale_c5_$n.json; done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_check.py 2>&1 | tail -3
