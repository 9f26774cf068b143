#!/bin/bash
# one 8-GPU box: the N=8 and N=4 bench lines (sharded master; config 4 and config 5 secondary blocks)
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "n8 rc=$?"
tail -c 1500 gpurun_out/r2_bench_n8.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
echo "n4 rc=$?"
tail -c 800 gpurun_out/r2_bench_n4.err
