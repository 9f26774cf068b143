#!/bin/bash
# the driver's scaling series on ONE 8-GPU box: N = 1, 2, 4, 8 back to back
set -x
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err; echo "n1 rc=$?"
for N in 2 4 8; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530+N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
  echo "n$N rc=$?"
done
tail -c 600 gpurun_out/r2_scale_n8.err
