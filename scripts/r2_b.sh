#!/bin/bash
# K2 after making the TMA / MMA roles warp-uniform: parity tests, then timings on the same box
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_scoring.py -x -q -k "tensorcore_scores or topk_matches_oracle" > gpurun_out/r2_qs_tests.txt 2>&1
tail -5 gpurun_out/r2_qs_tests.txt
echo "== classic pair kernel: 0 full | 64 one producer | 9 quarter MMAs no tmem reads | 1 no tmem reads" > gpurun_out/r2_qs_ablate.txt
QST_SCORE_QS=0 timeout 300 python profiles/ablate_k2.py 2 0,64,9,1,0 >> gpurun_out/r2_qs_ablate.txt 2>&1
echo "== query-stationary kernel: 0 full | 1 no TMEM reads | 16 thresholds at +inf | 129 no staging, no tmem reads" >> gpurun_out/r2_qs_ablate.txt
timeout 300 python profiles/ablate_k2.py 2 0,1,16,129,0 >> gpurun_out/r2_qs_ablate.txt 2>&1
cat gpurun_out/r2_qs_ablate.txt
