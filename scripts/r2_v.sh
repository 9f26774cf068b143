#!/bin/bash
# final loss kernel (limb reduction): plain run, then one ncu --set full capture of a warm launch
timeout 40 python profiles/run_loss.py > gpurun_out/r02_loss_final_plain.log 2>&1 || { tail -5 gpurun_out/r02_loss_final_plain.log; exit 1; }
tail -1 gpurun_out/r02_loss_final_plain.log
timeout 75 ncu --set full --clock-control none -k regex:quad_fused --launch-skip 8 -c 1 -o /tmp/prof_loss_final -f python profiles/run_loss.py > /tmp/ncu_loss_final.log 2>&1
tail -2 /tmp/ncu_loss_final.log
ncu -i /tmp/prof_loss_final.ncu-rep --page raw --csv > gpurun_out/r02_loss_final_raw.csv 2>/dev/null
ls -la gpurun_out/r02_loss_final_raw.csv
