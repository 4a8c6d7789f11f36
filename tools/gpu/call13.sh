#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/train_grad_diag.py 2 300,170 > gpurun_out/grad_diag_2.txt 2>&1; echo "diag exit $?"; cat gpurun_out/grad_diag_2.txt | tail -45
timeout 600 python tools/train_grad_diag.py 16 700,413 > gpurun_out/grad_diag_16.txt 2>&1; echo "diag exit $?"; cat gpurun_out/grad_diag_16.txt | tail -45
