#!/bin/bash
mkdir -p gpurun_out
RP_FMHA_V=1 python tools/kernel_bench.py fmha --iters 2 > gpurun_out/plain_v1.log 2>&1 &&
RP_FMHA_V=1 ncu --set full --clock-control none --import-source on -k regex:fmha_fwd -s 3 -c 1 -o gpurun_out/prof_r02_fmha_v1 python tools/kernel_bench.py fmha --iters 2 > gpurun_out/ncu_v1.log 2>&1
echo "v1 ncu exit $?"
RP_FMHA_V=3 python tools/kernel_bench.py fmha --iters 2 > gpurun_out/plain_v3.log 2>&1 &&
RP_FMHA_V=3 ncu --set full --clock-control none --import-source on -k regex:fmha_k64 -s 3 -c 1 -o gpurun_out/prof_r02_fmha_k64 python tools/kernel_bench.py fmha --iters 2 > gpurun_out/ncu_v3.log 2>&1
echo "v3 ncu exit $?"
python tools/comparator.py --skip-model --iters 2 > gpurun_out/plain_cmp.log 2>&1 &&
ncu --set full --clock-control none -k regex:sdpa -s 3 -c 1 -o gpurun_out/prof_r02_cudnn_sdpa python tools/comparator.py --skip-model --iters 2 > gpurun_out/ncu_cudnn.log 2>&1
echo "cudnn ncu exit $?"; tail -n 5 gpurun_out/ncu_cudnn.log
