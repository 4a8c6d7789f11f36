#!/bin/bash
# round 2, call 20 (2 GPUs): the two-device test, bench at N=2 under torchrun incl. the training step with its all-reduce
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -k "two_devices or two_gpu or second_device or devices" 2>&1 | tail -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_c20_n2.json 2> gpurun_out/bench_c20_n2.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_c20_n2.json').read().strip().splitlines()[-1])
    print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'e2e16', round(d['e2e_bf16']['value']))
    print(json.dumps(d['extra'].get('train_step'))[:900])
    print(json.dumps(d['extra'].get('config3_10k'))[-300:])
except Exception as e: print('parse failed', e); print(open('gpurun_out/bench_c20_n2.err').read()[-3000:])
PY
