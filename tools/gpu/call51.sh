#!/bin/bash
# round 2, call 51: fused attention backward, dropout variant: keep-bit words read four at a time: tests + A/B in the step
timeout 900 python -m pytest tests/test_gpu_dropout.py tests/test_gpu_train_kernels.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2; do for v in new head; do
  if [ $v = new ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_$v.so; fi
  timeout 600 python tools/train_bench.py --B 16 --dropout 0.1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v train B16 dropout: ms', round(d['ms_per_step'],2), 'bwd_fmha', d['kernel_classes_ms']['bwd_fmha'])"
done; done
