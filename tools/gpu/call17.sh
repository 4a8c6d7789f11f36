#!/bin/bash
# round 2, call 17: head-dot epilogue tests + bench; attention-backward kernel split; concat_cast ncu
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "gemm or head" -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_comparator.py -q -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_c17.json 2> gpurun_out/bench_c17.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_c17.json').read().strip().splitlines()[-1])
    print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'fmha frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])
    for k,v in d['kernels'].items(): print(' ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
except Exception as e: print('parse failed', e); print(open('gpurun_out/bench_c17.err').read()[-2000:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fmha_bwd|attn_bwd|fmha_fwd" --csv --log-file gpurun_out/r02_fmha_bwd_launches.csv python tools/kernel_bench.py fmhabwd --iters 1 > /dev/null 2>&1; echo "ncu exit $?"
python tools/summarize_ncu.py launches gpurun_out/r02_fmha_bwd_launches.csv
timeout 300 ncu --set full --clock-control none -k regex:"concat_cast" -c 2 -o gpurun_out/r02_concat python tools/kernel_bench.py rowwise --iters 1 > /dev/null 2>&1; echo "ncu exit $?"
