#!/bin/bash
# round 2, call 38: fused attention backward with 16 softmax warps (and the dQ MMA's wait moved behind dV / dK): tests, A/B
# against the 8-warp build (ab/lib_fused8.so), ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_dropout.py tests/test_gpu_train_step.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
for rep in 1 2; do for v in new fused8; do
  if [ $v = new ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_$v.so; fi
  for B in 16 32; do echo -n "$v B=$B: "; timeout 300 python tools/kernel_bench.py fmhabwd --iters 20 --B $B 2>&1 | tr -d '\n '; echo; done
done; done
unset RP_LIB_PATH
for d in 0.0 0.1; do timeout 600 python tools/train_bench.py --B 16 --dropout $d > gpurun_out/train_bench_c38_$d.json 2> gpurun_out/train_bench_c38_$d.err; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/train_bench_c38_$d.json').read().strip().splitlines()[-1])
    print('dropout $d: ms/step', round(d['ms_per_step'],2), 'bwd_fmha', d['kernel_classes_ms'].get('bwd_fmha'))
except Exception as e: print('parse failed', e); print(open('gpurun_out/train_bench_c38_$d.err').read()[-1500:])
PY
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fmha_bwd_fused" -s 2 -c 1 -o gpurun_out/r02_bwd_fused16 python tools/kernel_bench.py fmhabwd --iters 1 --B 16 > gpurun_out/ncu_bwd_fused.log 2>&1; echo "ncu exit $?"
