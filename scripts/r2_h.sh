#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_h_tests.txt 2>&1
tail -15 gpurun_out/r2_h_tests.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r2_bench_n1.err
