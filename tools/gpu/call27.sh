#!/bin/bash
# round 2, call 27: eight epilogue warps for the K <= 512 bf16-epilogue GEMMs (QKV, FF1, heads): tests, A/B alone, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "gemm" 2>&1 | tail -2
for e in 0 1; do echo "== RP_GEMM_EW8=$e"; RP_GEMM_EW8=$e timeout 300 python tools/kernel_bench.py gemm --iters 20 2>&1 | tr -d '\n ' | grep -o '"gemm_qkv[^}]*}\|"gemm_ff1[^}]*}'; echo; done
for e in 0 1; do
RP_GEMM_EW8=$e timeout 900 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_c27_$e.json 2> gpurun_out/bench_c27_$e.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c27_$e.json').read().strip().splitlines()[-1])
    print('EW8=$e value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'fmha frac', round(d['roofline']['frac'],3))
    for k in ('gemm_qkv','gemm_ff1','gemm_head','fmha','gemm_ff2','gemm_out'): print(' ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in d['kernels'][k].items()})
except Exception as e: print('parse failed', e); print(open('gpurun_out/bench_c27_$e.err').read()[-2000:])
PY
done
