#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharded_emulated.py -x -q > gpurun_out/r2_n8_tests.txt 2>&1
tail -3 gpurun_out/r2_n8_tests.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8b.json 2> gpurun_out/r2_bench_n8b.err
echo "n8 rc=$?"
tail -c 800 gpurun_out/r2_bench_n8b.err
