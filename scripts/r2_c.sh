#!/bin/bash
mkdir -p gpurun_out
timeout 300 profiles/build/ubench_umma > gpurun_out/r2_ubench_umma.txt 2>&1
cat gpurun_out/r2_ubench_umma.txt
