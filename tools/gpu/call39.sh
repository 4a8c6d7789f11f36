#!/bin/bash
# round 2, call 39: full validation on a fresh box: every GPU test, smoke(), the default bench (both arms), train bench B=32
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/bench_c39.json 2> gpurun_out/bench_c39.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_c39.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c39_ref.json 2> gpurun_out/bench_c39_ref.err; echo "ref exit $?"; tail -1 gpurun_out/bench_c39_ref.json | cut -c1-600
for B in 16 32; do for d in 0.0 0.1; do timeout 600 python tools/train_bench.py --B $B --dropout $d > gpurun_out/train_bench_c39_${B}_$d.json 2> gpurun_out/train_bench_c39_${B}_$d.err; done; done
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c39.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'], d['clocks'])
print('cpu_baseline', d['cpu_baseline'])
for k,v in (d.get('extra') or {}).items(): print(k, json.dumps(v)[:500])
for B in (16,32):
    for dr in ('0.0','0.1'):
        try:
            t=json.loads(open(f'gpurun_out/train_bench_c39_{B}_{dr}.json').read().strip().splitlines()[-1])
            print('train B',B,'dropout',dr,'ms',round(t['ms_per_step'],2),'videos/s',round(t['videos_per_s']),'TF',round(t['model_tflops_per_gpu']), {k:v['ms'] for k,v in list(t['kernel_classes_ms'].items())[:5]})
        except Exception as e: print('train parse failed',B,dr,e)
PY
