#!/bin/bash
# round 2, call 43: fused attention backward: statistics preloaded before the S^T barrier wait (0 / 4 / 8 float4 pairs)
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_dropout.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2; do for v in p0 main p8; do
  if [ $v = main ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_$v.so; fi
  for B in 16; do echo -n "$v B=$B: "; timeout 300 python tools/kernel_bench.py fmhabwd --iters 30 --B $B 2>&1 | tr -d '\n ' | sed 's/.*fmha_bwd"://'; echo; done
done; done
