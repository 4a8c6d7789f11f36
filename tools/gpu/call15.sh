#!/bin/bash
# round 2, call 15: timings of the backward kernels and the training step, default bench wall time, launch list
mkdir -p gpurun_out
timeout 300 python tools/kernel_bench.py fmhabwd rowwise --iters 10 2>&1 | tr -d '\n '; echo
timeout 600 python tools/train_bench.py --B 16 --T 1801 --steps 3 > gpurun_out/train_bench_b16.json 2> gpurun_out/train_bench_b16.err; echo "train bench exit $?"; cat gpurun_out/train_bench_b16.json; tail -3 gpurun_out/train_bench_b16.err
SECONDS=0
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default bench exit $? in ${SECONDS}s"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
    print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'fmha frac', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
    print(json.dumps(d['extra'])[:3000])
except Exception as e: print('parse failed', e); print(open('gpurun_out/bench_default.err').read()[-2000:])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
