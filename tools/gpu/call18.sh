#!/bin/bash
# round 2, call 18: paired chunk walk in the 4-warp residual epilogue: tests, A/B, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "gemm" -p no:cacheprovider 2>&1 | tail -3
for pz in 0 1; do
  echo "== RP_EPI_PAIRED=$pz"
  RP_EPI_PAIRED=$pz timeout 300 python tools/kernel_bench.py gemmln gemm --iters 20 2>&1 | tr -d '\n ' | grep -o '"gemm_out(epi3)[^}]*}\|"gemm_ff2(epi3)[^}]*}\|"gemm_out+ln[^}]*}\|"gemm_ff2+ln[^}]*}'; echo
done
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_comparator.py tests/test_gpu_train_step.py -q -p no:cacheprovider 2>&1 | tail -3
for pz in 0 1; do
RP_EPI_PAIRED=$pz timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_c18_$pz.json 2> gpurun_out/bench_c18.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c18_$pz.json').read().strip().splitlines()[-1])
    print('paired=$pz value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'fmha frac', round(d['roofline']['frac'],3))
    for k in ('fmha','gemm_out','gemm_ff2','gemm_ff1','gemm_qkv'): print(' ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in d['kernels'][k].items()})
except Exception as e: print('parse failed', e); print(open('gpurun_out/bench_c18.err').read()[-2000:])
PY
done
