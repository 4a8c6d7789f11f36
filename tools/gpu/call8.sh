#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/test_$name.log 2>&1; echo "== $name: exit $?"; tail -n 15 gpurun_out/test_$name.log; }
run model tests/test_gpu_model.py
run comparator tests/test_gpu_comparator.py
run kernels tests/test_gpu_kernels.py
timeout 600 python tools/latency_bs1.py > gpurun_out/latency_bs1.log 2>&1; echo "latency exit $?"; tail -n 6 gpurun_out/latency_bs1.log
timeout 900 python bench.py > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench_r02_b.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])
print(json.dumps(d['extra'])[:1500])
print({k:(round(v['ms_per_step'],3), round(v.get('tflops',0) or v.get('gbs',0),1)) for k,v in d['kernels'].items()})
PY
