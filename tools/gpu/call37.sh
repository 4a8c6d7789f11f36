#!/bin/bash
# round 2, call 37: fused attention backward: the extended parity tests (both modes), then ncu of the fused kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_dropout.py tests/test_gpu_train_step.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -v "^  \|Warning\|warnings.warn" | tail -25
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fmha_bwd_fused" -s 2 -c 1 -o gpurun_out/r02_bwd_fused python tools/kernel_bench.py fmhabwd --iters 1 --B 16 > gpurun_out/ncu_bwd_fused.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fmha_bwd|attn_bwd|Memset|memset" -c 12 --csv --log-file gpurun_out/r02_fmha_bwd_fused_launches.csv python tools/kernel_bench.py fmhabwd --iters 1 --B 16 > /dev/null 2>&1; echo "ncu exit $?"
