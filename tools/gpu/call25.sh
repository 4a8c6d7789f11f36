#!/bin/bash
# round 2, call 25: deeper TMA rings in the attention backward kernels: tests, kernel alone, training bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropout.py tests/test_gpu_train_kernels.py tests/test_gpu_train_step.py -q -p no:cacheprovider -x 2>&1 | tail -2
timeout 300 python tools/kernel_bench.py fmhabwd --iters 10 --B 16 2>&1 | tr -d '\n '; echo
for d in 0.1 0.0; do
timeout 600 python tools/train_bench.py --B 16 --steps 4 --dropout $d > gpurun_out/train_bench_c25_$d.json 2> gpurun_out/train_bench_c25_$d.err; echo "train bench $d exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/train_bench_c25_$d.json').read().strip().splitlines()[-1])
    print(round(d['ms_per_step'],2),'ms', round(d['forward_ms'],2), 'fwd ms', d['loss_first_last'], json.dumps(d['kernel_classes_ms']))
except Exception as e: print('parse failed', e); print(open('gpurun_out/train_bench_c25_$d.err').read()[-2000:])
PY
done
