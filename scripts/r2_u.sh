#!/bin/bash
# loss kernel: limb reduction (default) vs fence + ticket (QST_LOSS_REDUCE=ticket); loss parity tests first
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_loss.py -x -q > gpurun_out/r2u_loss_tests.txt 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2u_loss_tests.txt | cut -c1-250
timeout 120 python profiles/loss_probe.py > gpurun_out/r2u_probe_limbs.txt 2>&1; echo "limbs rc=$?"; cat gpurun_out/r2u_probe_limbs.txt | cut -c1-200
QST_LOSS_REDUCE=ticket timeout 120 python profiles/loss_probe.py > gpurun_out/r2u_probe_ticket.txt 2>&1; echo "ticket rc=$?"; cat gpurun_out/r2u_probe_ticket.txt | cut -c1-200
