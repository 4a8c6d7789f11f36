#!/bin/bash
# round 2, call 59: end-to-end step fed from write-combined pinned host memory vs torch pinned memory
for rep in 1 2 3; do for wc in 0 1; do
  RP_BENCH_WC=$wc timeout 900 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('wc=$wc value', round(d['value']), 'e2e', round(e['value']), 'e2e ms', round(e['ms_per_step'],3), 'serial ms', round(e['serial_inference__ms_per_step'],2), d['clocks']['sm_mhz'])"
done; done
