#!/bin/bash
# round 2, call 56: L2 zig-zag A/B, five alternating pairs of 40-step runs
for rep in 1 2 3 4 5; do for z in 0 1; do
  RP_ZIGZAG=$z timeout 900 python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('zigzag=$z ms', round(d['ms_per_step'],3), 'e2e ms', round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'])"
done; done
