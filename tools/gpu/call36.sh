#!/bin/bash
# round 2, call 36: fused attention backward (dK, dV, dQ in one kernel; dQ through the TMA reduce path): parity tests,
# then A/B against the two deterministic kernels (RP_FMHA_BWD_FUSED=0), alone and in the training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_dropout.py tests/test_gpu_train_step.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -8
for rep in 1 2; do for f in 1 0; do
  for B in 16 32; do echo -n "fused=$f B=$B: "; RP_FMHA_BWD_FUSED=$f timeout 300 python tools/kernel_bench.py fmhabwd --iters 20 --B $B 2>&1 | tr -d '\n '; echo; done
done; done
for f in 1 0; do
  for d in 0.0 0.1; do RP_FMHA_BWD_FUSED=$f timeout 600 python tools/train_bench.py --B 16 --dropout $d > gpurun_out/train_bench_c36_${f}_$d.json 2> gpurun_out/train_bench_c36_${f}_$d.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/train_bench_c36_${f}_$d.json').read().strip().splitlines()[-1])
    print('fused=$f dropout $d: ms/step', round(d['ms_per_step'],2), 'bwd_fmha', d['kernel_classes_ms'].get('bwd_fmha'), 'loss', d['loss_first_last'])
except Exception as e: print('parse failed', e); print(open('gpurun_out/train_bench_c36_${f}_$d.err').read()[-1500:])
PY
  done
done
