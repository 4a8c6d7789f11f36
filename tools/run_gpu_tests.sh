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
run comparator tests/test_gpu_comparator.py
