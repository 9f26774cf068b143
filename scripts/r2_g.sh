#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sharded_emulated.py -x -q > gpurun_out/r2_g_tests.txt 2>&1
tail -30 gpurun_out/r2_g_tests.txt
