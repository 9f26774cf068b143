#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/k3_probe.py 128,144,160,176,192,224 > gpurun_out/r2_k3_kprime.txt 2>&1
cat gpurun_out/r2_k3_kprime.txt
