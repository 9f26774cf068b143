#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scoring.py -x -q > gpurun_out/r2_e_tests.txt 2>&1
tail -3 gpurun_out/r2_e_tests.txt
timeout 900 python profiles/k2_ab.py "classic:QST_SCORE_QS=0,QST_CHUNKMAX=1" "qs:QST_SCORE_QS=1,QST_CHUNKMAX=1" "classic_noepi:QST_SCORE_QS=0,QST_SCORE_DEBUG=1" "qs_noepi:QST_SCORE_QS=1,QST_SCORE_DEBUG=1" --blocks 3 --launches 40 > gpurun_out/r2_k2_ab3.txt 2>&1
cat gpurun_out/r2_k2_ab3.txt
