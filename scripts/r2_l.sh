#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/loss_probe.py > gpurun_out/r2_loss_probe.txt 2>&1
cat gpurun_out/r2_loss_probe.txt
timeout 900 python -m pytest tests -x -q -m gpu -k "quad or loss" 2>&1 | tail -5
