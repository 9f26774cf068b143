#!/bin/bash
# why is the query-stationary K2 slow?  ablations + one ncu capture
set -x
mkdir -p gpurun_out
echo "== QS: 0 full | 1 no tmem reads | 129 +no staging | 257 no MMAs | 513 all-TS | 1025 all-SS | 9 quarter MMAs" > gpurun_out/r2_qs_ablate2.txt
timeout 600 python profiles/ablate_k2.py 2 0,1,129,257,513,1025,9 >> gpurun_out/r2_qs_ablate2.txt 2>&1
cat gpurun_out/r2_qs_ablate2.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_select -c 1 -o gpurun_out/prof_k2_qs_v2 -f python profiles/run_k2_qs.py > gpurun_out/ncu_qs.log 2>&1
tail -3 gpurun_out/ncu_qs.log
