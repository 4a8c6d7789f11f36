#!/bin/bash
# round 2, call 60: final ncu capture of the fused attention backward (+ its launch list)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fmha_bwd_fused" -s 2 -c 1 -o gpurun_out/r02_bwd_fused_final python tools/kernel_bench.py fmhabwd --iters 1 --B 16 > gpurun_out/ncu_bwd_fused.log 2>&1; echo "ncu exit $?"
for B in 16 32; do echo -n "B=$B: "; timeout 300 python tools/kernel_bench.py fmhabwd --iters 30 --B $B 2>&1 | tr -d '\n ' | sed 's/.*fmha_bwd"://'; echo; done
