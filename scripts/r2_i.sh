#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python profiles/loss_api_probe.py > gpurun_out/r2_loss_api.txt 2>&1
head -40 gpurun_out/r2_loss_api.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_i_tests.txt 2>&1
tail -5 gpurun_out/r2_i_tests.txt
