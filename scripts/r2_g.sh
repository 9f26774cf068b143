#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_g_tests.txt 2>&1
tail -4 gpurun_out/r2_g_tests.txt
QST_K3_FIRST=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_k3off.json 2> gpurun_out/r2_bench_k3off.err; echo "rc=$?"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_k3on.json 2> gpurun_out/r2_bench_k3on.err; echo "rc=$?"
QST_K3_FIRST=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_k3off2.json 2> /dev/null; echo "rc=$?"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_k3on2.json 2> /dev/null; echo "rc=$?"
