#!/bin/bash
# round 2, call 24: attention backward alone (B=16) + ncu --set full of the dQ and dK/dV kernels
mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py fmhabwd --iters 10 --B 16 2>&1 | tail -1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fmha_bwd_dkdv" -s 1 -c 1 -o gpurun_out/r02_bwd_dkdv python tools/kernel_bench.py fmhabwd --iters 1 --B 16 > gpurun_out/ncu_dkdv.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fmha_bwd_dq" -s 1 -c 1 -o gpurun_out/r02_bwd_dq python tools/kernel_bench.py fmhabwd --iters 1 --B 16 > gpurun_out/ncu_dq.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/*.ncu-rep
