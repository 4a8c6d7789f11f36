#!/bin/bash
# 8-GPU box: topology probe, then the bench at N=8 and N=4 (value, e2e, config 3)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 300 python tools/topo_probe.py --mb 512 --reps 6 > gpurun_out/topo_probe.json 2> gpurun_out/topo_probe.err; echo "probe exit $?"; head -c 1500 gpurun_out/topo_probe.json
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_r02_n$n.json 2> gpurun_out/bench_r02_n$n.err; echo "bench n=$n exit $?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r02_n$n.json').read().strip().splitlines()[-1])
    print($n, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e_bf16', round(d['e2e']['bf16_feature_rows']['value']), 'numa', d['e2e'].get('host_numa_node'), json.dumps(d['extra'].get('config3_10k',{}))[-400:])
except Exception as e: print('parse failed', e)
PY
done
