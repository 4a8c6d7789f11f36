#!/bin/bash
# round 2, call 61: vectorised bias column sums: training tests + step time
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_step.py tests/test_gpu_dropout.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
for d in 0.0 0.1; do timeout 600 python tools/train_bench.py --B 16 --dropout $d 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernel_classes_ms']; print('dropout $d: ms', round(d['ms_per_step'],2), {n:k[n]['ms'] for n in ('bwd_bias_colsum','bwd_relu','bwd_layernorm','bwd_fmha')})"
done
