#!/bin/bash
# round 2, call 44: fused attention backward with full statistic preload: tests + ncu
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_dropout.py tests/test_gpu_train_step.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
for B in 16 32; do echo -n "B=$B: "; timeout 300 python tools/kernel_bench.py fmhabwd --iters 30 --B $B 2>&1 | tr -d '\n ' | sed 's/.*fmha_bwd"://'; echo; done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fmha_bwd_fused" -s 2 -c 1 -o gpurun_out/r02_bwd_fused_v2 python tools/kernel_bench.py fmhabwd --iters 1 --B 16 > gpurun_out/ncu_bwd_fused.log 2>&1; echo "ncu exit $?"
