#!/bin/bash
# round 2, call 48: after ranking-by-length in every multi-video forward: model / kernel tests, default bench without extras
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py tests/test_gpu_comparator.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c48.json 2> gpurun_out/bench_c48.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c48.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), d['clocks'])
for k in ('config3_10k','ragged_unsorted_batch'):
    print(k, json.dumps(d['extra'][k])[:700])
PY
