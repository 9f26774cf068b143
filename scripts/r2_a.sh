#!/bin/bash
# round-2 baseline: K2 ablations (is the main loop L2-feed-bound?) + the bench line of the r01 kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
python profiles/ablate_k2.py 2 0,1,9,5,13,0 > gpurun_out/r2_ablate.txt 2>&1
cat gpurun_out/r2_ablate.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_base.json 2> gpurun_out/r2_bench_base.err
tail -c 3000 gpurun_out/r2_bench_base.json
