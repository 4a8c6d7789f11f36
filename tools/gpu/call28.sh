#!/bin/bash
# round 2, call 28: persistent attention forward: parity tests, A/B alone and in the step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "fmha" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_comparator.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
for e in 0 1; do echo "== RP_FMHA_PERSIST=$e"; for T in 1801 1792 8192; do B=32; if [ $T = 8192 ]; then B=4; fi; RP_FMHA_PERSIST=$e timeout 300 python tools/kernel_bench.py fmha --iters 20 --T $T --B $B 2>&1 | tr -d '\n '; echo " T=$T"; done; done
for e in 0 1; do
RP_FMHA_PERSIST=$e timeout 900 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_c28_$e.json 2> gpurun_out/bench_c28_$e.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c28_$e.json').read().strip().splitlines()[-1])
    print('PERSIST=$e value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'fmha frac', round(d['roofline']['frac'],3), 'fmha', d['kernels']['fmha'])
except Exception as e: print('parse failed', e); print(open('gpurun_out/bench_c28_$e.err').read()[-2000:])
PY
done
