#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python profiles/k2_ab.py "s46:QST_SCORE_QS=1" "s37:QST_STRIPES=37" "s74:QST_STRIPES=74" --blocks 3 --launches 40 > gpurun_out/r2_stripes2.txt 2>&1
QST_STRIPES=37 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:score_select -c 1 python profiles/run_k2_qs.py 2>&1 | grep -E "dram__bytes|gpu__time|qs " >> gpurun_out/r2_stripes2.txt
cat gpurun_out/r2_stripes2.txt
