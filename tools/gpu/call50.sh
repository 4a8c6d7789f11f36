#!/bin/bash
# round 2, call 50 (8 GPUs): the driver's scaling command on the final tree
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_c50_n8.json 2> gpurun_out/bench_c50_n8.err; echo "exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c50_n8.json').read().strip().splitlines()[-1])
print('n', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'e2e_bf16', round((d.get('e2e_bf16') or d['e2e'].get('bf16_feature_rows',{})).get('value',0)))
for k,v in (d.get('extra') or {}).items(): print(k, json.dumps(v)[:500])
PY
tail -3 gpurun_out/bench_c50_n8.err
