#!/bin/bash
# round 2, call 55: L2 zig-zag (every kernel of a layer walks the rows opposite to its producer): parity + A/B in the step
mkdir -p gpurun_out
RP_ZIGZAG=1 timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -2
for rep in 1 2 3; do for z in 0 1; do
  RP_ZIGZAG=$z timeout 900 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; print('zigzag=$z value', round(d['value']), 'ms', round(d['ms_per_step'],3), {n: round(k[n]['ms_per_step'],3) for n in ('gemm_qkv','fmha','gemm_out','gemm_ff1','gemm_ff2')}, d['clocks']['sm_mhz'])"
done; done
