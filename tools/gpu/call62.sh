#!/bin/bash
# round 2, call 62 (4 GPUs): the driver's scaling command on the final tree
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_c62_n4.json 2> gpurun_out/bench_c62_n4.err; echo "exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c62_n4.json').read().strip().splitlines()[-1])
print('n', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'bf16', round(d['e2e'].get('bf16_feature_rows',{}).get('value',0)))
ex=d.get('extra') or {}
print('config3', ex.get('config3_10k',{}).get('fp32_rows'), ex.get('config3_10k',{}).get('bf16_rows'))
print('train', ex.get('train_step',{}).get('ms_per_step'), ex.get('train_step_dropout_off',{}).get('ms_per_step'))
PY
