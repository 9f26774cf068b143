#!/bin/bash
# evidence for profiles/: plain bench line, then (each only after its plain command exited 0) the ncu launch
# list of the same command and ncu captures of the top kernels.  Reports are exported to CSV pages on the box
# (raw + source) and removed: gpurun only copies back 64 MiB.
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err || exit 1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err || exit 1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /tmp/ncu_launches.log 2>&1
python profiles/run_k2_qs.py 2 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_select -c 1 -o /tmp/prof_k2_qs_r02 -f python profiles/run_k2_qs.py > /tmp/ncu_qs_r02.log 2>&1
tail -2 /tmp/ncu_qs_r02.log
ncu -i /tmp/prof_k2_qs_r02.ncu-rep --page raw --csv > gpurun_out/r02_k2_qs_raw.csv 2>/dev/null
ncu -i /tmp/prof_k2_qs_r02.ncu-rep --page source --csv > gpurun_out/r02_k2_qs_source.csv 2>/dev/null
python profiles/run_loss.py || exit 1
timeout 600 ncu --set full --clock-control none -k regex:quad_fused -c 1 -o /tmp/prof_loss_r02 -f python profiles/run_loss.py > /tmp/ncu_loss_r02.log 2>&1
tail -2 /tmp/ncu_loss_r02.log
ncu -i /tmp/prof_loss_r02.ncu-rep --page raw --csv > gpurun_out/r02_loss_raw.csv 2>/dev/null
ls -la gpurun_out /tmp/*.ncu-rep
du -sh gpurun_out
