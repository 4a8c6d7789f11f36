#!/bin/bash
# Runs the GPU test groups in separate processes (a trapped kernel poisons its CUDA context, so one
# failing group must not hide the others).  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, pytest args...
  local name=$1; shift
  timeout 600 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/test_$name.log 2>&1
  echo "== $name: exit $?"; tail -n 25 gpurun_out/test_$name.log
}
run gemm tests/test_gpu_kernels.py -k "gemm"
run rowwise tests/test_gpu_kernels.py -k "concat or layernorm or head_out"
run fmha tests/test_gpu_kernels.py -k "fmha"
run model tests/test_gpu_model.py
RP_LN_IN_GEMM=0 run model_ln_standalone tests/test_gpu_model.py -k "forward or inference or full_size"
RP_FMHA_NQ=2 run fmha_nq2 tests/test_gpu_kernels.py -k fmha
RP_GEMM_CG=1 run gemm_cg1 tests/test_gpu_kernels.py -k gemm
