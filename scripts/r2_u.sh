#!/bin/bash
# loss kernel: parity tests, then the reduction variants side by side in one process (profiles/loss_probe.py)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_loss.py -x -q > gpurun_out/r2u_loss_tests.txt 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2u_loss_tests.txt | cut -c1-250
QST_PROBE_MODES=default,ticket timeout 200 python profiles/loss_probe.py > gpurun_out/r2u_probe_modes.txt 2>&1; echo "probe rc=$?"; cat gpurun_out/r2u_probe_modes.txt | cut -c1-200
