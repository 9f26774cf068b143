#!/bin/bash
set -x
mkdir -p gpurun_out
QST_SCORE_QS=0 QST_CHUNKMAX=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_select -c 1 -o gpurun_out/prof_k2_classic_lean -f python profiles/run_k2_qs.py > gpurun_out/ncu_classic.log 2>&1
tail -2 gpurun_out/ncu_classic.log
QST_SCORE_QS=1 QST_CHUNKMAX=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_select -c 1 -o gpurun_out/prof_k2_qs_lean -f python profiles/run_k2_qs.py > gpurun_out/ncu_qs.log 2>&1
tail -2 gpurun_out/ncu_qs.log
