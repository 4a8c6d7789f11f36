#!/bin/bash
# round 2, call 32: ncu at T=8192, B=4 (64 key tiles per item: steady state) of the persistent and the shipped attention kernel
mkdir -p gpurun_out
RP_FMHA_PERSIST=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fmha_fwd_persist" -s 2 -c 1 -o gpurun_out/r02_fmha_persist8k python tools/kernel_bench.py fmha --iters 1 --T 8192 --B 4 > gpurun_out/ncu_persist.log 2>&1; echo "ncu exit $?"
RP_FMHA_PERSIST=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fmha_fwd_kernel" -s 2 -c 1 -o gpurun_out/r02_fmha_shipped8k python tools/kernel_bench.py fmha --iters 1 --T 8192 --B 4 > gpurun_out/ncu_shipped.log 2>&1; echo "ncu exit $?"
