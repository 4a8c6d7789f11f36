#!/bin/bash
# round 2, call 42: fused attention backward with per-warp statistic staging (no named barrier per tile): tests, A/B
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_dropout.py tests/test_gpu_train_step.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
for rep in 1 2; do for v in new head; do
  if [ $v = new ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_$v.so; fi
  for B in 16 32; do echo -n "$v B=$B: "; timeout 300 python tools/kernel_bench.py fmhabwd --iters 20 --B $B 2>&1 | tr -d '\n '; echo; done
done; done
for v in new head; do
  if [ $v = new ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_$v.so; fi
  timeout 600 python tools/train_bench.py --B 16 --dropout 0.1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v train B16 dropout: ms', round(d['ms_per_step'],2), 'bwd_fmha', d['kernel_classes_ms']['bwd_fmha'])"
done
