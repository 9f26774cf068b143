#!/bin/bash
# 2 GPUs: real-NCCL test of the sharded path (incl. the C-ABI communicator, copy-engine query gather) + the N=2 bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded_emulated.py -x -q > gpurun_out/r2_n2_tests.txt 2>&1
tail -15 gpurun_out/r2_n2_tests.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench rc=$?"
tail -c 2000 gpurun_out/r2_bench_n2.err
