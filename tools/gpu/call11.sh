#!/bin/bash
# round 2, call 11: backward-kernel unit tests; residual L2-prefetch A/B; ncu of the LayerNorm-fused out-proj GEMM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py -q -x -p no:cacheprovider > gpurun_out/test_train_kernels.log 2>&1; echo "train kernels exit $?"; tail -n 30 gpurun_out/test_train_kernels.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "gemm or fmha or head_out or layernorm" -p no:cacheprovider 2>&1 | tail -3
for pf in 0 1; do
  echo "== RP_RESID_PREFETCH=$pf"
  RP_RESID_PREFETCH=$pf timeout 300 python tools/kernel_bench.py gemmln gemm rowwise ln --iters 20 2>&1 | tr -d '\n '; echo
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"gemm_bf16_kernel" -c 1 -o gpurun_out/r02_gemm_outln python tools/kernel_bench.py gemmln --iters 1 > gpurun_out/ncu_outln.log 2>&1; echo "ncu exit $?"
