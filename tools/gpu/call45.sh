#!/bin/bash
# round 2, call 45: fused attention backward: S^T / dP^T issued two tiles ahead of the gradient MMAs: tests, A/B vs HEAD
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_dropout.py tests/test_gpu_train_step.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2; do for v in new head; do
  if [ $v = new ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_$v.so; fi
  for B in 16 32; do echo -n "$v B=$B: "; timeout 300 python tools/kernel_bench.py fmhabwd --iters 30 --B $B 2>&1 | tr -d '\n ' | sed 's/.*fmha_bwd"://'; echo; done
done; done
