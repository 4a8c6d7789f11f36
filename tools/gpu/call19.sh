#!/bin/bash
# round 2, call 19: padding skip (RowMap + attention query-tile skip): full GPU suite, bench with extras
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh 2>&1 | grep -E "^== |passed|failed|error" | head -40
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_train_kernels.py tests/test_train_pieces.py -q -p no:cacheprovider 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/bench_c19.json 2> gpurun_out/bench_c19.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_c19.json').read().strip().splitlines()[-1])
    print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'fmha frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])
    print(json.dumps(d['extra'].get('ragged_unsorted_batch')))
    print(json.dumps(d['extra'].get('config3_10k'))[-300:])
    print(json.dumps(d['extra'].get('train_step'))[:300])
except Exception as e: print('parse failed', e); print(open('gpurun_out/bench_c19.err').read()[-2000:])
PY
