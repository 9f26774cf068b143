#!/bin/bash
mkdir -p gpurun_out
timeout 600 python profiles/overlap_probe.py > gpurun_out/r2_overlap.txt 2>&1
cat gpurun_out/r2_overlap.txt
