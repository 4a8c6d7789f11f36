#!/bin/bash
# round 2, call 40: 2-GPU bench (incl. the training-step extra with the fused attention backward and the gradient all-reduce)
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_c40_n2.json 2> gpurun_out/bench_c40_n2.err; echo "exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c40_n2.json').read().strip().splitlines()[-1])
print('n', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
for k,v in (d.get('extra') or {}).items(): print(k, json.dumps(v)[:400])
PY
timeout 300 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -k "two_dev or two_device" 2>&1 | tail -2
