#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded_emulated.py -x -q -m gpu 2>&1 | tail -40 > gpurun_out/r2_t.txt
cat gpurun_out/r2_t.txt
