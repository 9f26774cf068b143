#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_g_tests.txt 2>&1
tail -6 gpurun_out/r2_g_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.txt 2>&1; tail -2 gpurun_out/r2_smoke.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_chk.json 2> gpurun_out/r2_bench_chk.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r2_bench_ref.json | cut -c1-600
