#!/bin/bash
# round 2, call 35: dK/dV kernel with decoupled softmax warps + TMEM prefetch: parity tests, then A/B against the
# committed kernel (ab/lib_oldbwd.so) alone (B = 16 and 32) and in the training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_dropout.py tests/test_gpu_train_step.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
for rep in 1 2; do for v in new old; do
  if [ $v = new ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_oldbwd.so; fi
  for B in 16 32; do echo -n "$v B=$B: "; timeout 300 python tools/kernel_bench.py fmhabwd --iters 20 --B $B 2>&1 | tail -1; done
done; done
for v in new old; do
  if [ $v = new ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_oldbwd.so; fi
  for d in 0.0 0.1; do timeout 600 python tools/train_bench.py --B 16 --dropout $d > gpurun_out/train_bench_c35_${v}_$d.json 2> gpurun_out/train_bench_c35_${v}_$d.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/train_bench_c35_${v}_$d.json').read().strip().splitlines()[-1])
    print('$v dropout $d: ms/step', round(d['ms_per_step'],2), 'bwd_fmha', d['kernel_classes_ms'].get('bwd_fmha'))
except Exception as e: print('parse failed', e); print(open('gpurun_out/train_bench_c35_${v}_$d.err').read()[-1500:])
PY
  done
done
