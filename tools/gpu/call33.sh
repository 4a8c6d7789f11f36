#!/bin/bash
# round 2, call 33: A/B of (a) the K ring depth of the attention forward (2 vs 3 stages), (b) a suspend-time hint on
# mbarrier.try_wait (fmha only / every kernel), alone and in the step.  The persistent kernel is switched off.
mkdir -p gpurun_out
export RP_FMHA_PERSIST=0
for rep in 1 2; do
for v in main k2 hint; do
  if [ $v = main ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_$v.so; fi
  echo "== $v (rep $rep)"
  for T in 1801 1792 8192; do B=32; if [ $T = 8192 ]; then B=4; fi; timeout 300 python tools/kernel_bench.py fmha --iters 20 --T $T --B $B 2>&1 | tr -d '\n '; echo " T=$T"; done
done; done
for v in main hintall main hintall; do
  if [ $v = main ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_$v.so; fi
  timeout 900 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_c33_$v.json 2> gpurun_out/bench_c33_$v.err; echo "bench exit $?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c33_$v.json').read().strip().splitlines()[-1])
    k=d['kernels']
    print('$v value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'fmha frac', round(d['roofline']['frac'],3), {n:round(k[n]['ms_per_step'],3) for n in k}, d['clocks'])
except Exception as e: print('parse failed', e); print(open('gpurun_out/bench_c33_$v.err').read()[-2000:])
PY
done
