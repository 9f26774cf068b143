#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scoring.py tests/test_gpu_sharded_emulated.py -x -q > gpurun_out/r2_e_tests.txt 2>&1
tail -3 gpurun_out/r2_e_tests.txt
timeout 900 python profiles/k2_ab.py "grouped:QST_SCORE_QS=1" "flat:QST_SCORE_DEBUG=2048" "grouped_cm2:QST_CHUNKMAX=2" "classic_grouped:QST_SCORE_QS=0" --blocks 3 --launches 40 > gpurun_out/r2_k2_ab4.txt 2>&1
cat gpurun_out/r2_k2_ab4.txt
