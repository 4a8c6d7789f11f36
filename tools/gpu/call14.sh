#!/bin/bash
# round 2, call 14: 8-warp LayerNorm-fused epilogue (K <= 512): tests, A/B; training step parity tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "gemm" -p no:cacheprovider 2>&1 | tail -5
for w in 4 8; do
  echo "== RP_EPI_WARPS=$w"
  RP_EPI_WARPS=$w timeout 300 python tools/kernel_bench.py gemmln --iters 20 2>&1 | tr -d '\n '; echo
done
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_comparator.py -q -p no:cacheprovider 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_train_kernels.py tests/test_train_pieces.py -q -s -p no:cacheprovider > gpurun_out/test_train_step.log 2>&1; echo "train tests exit $?"; grep -E "passed|failed|largest|loss per" gpurun_out/test_train_step.log | cut -c1-700
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_c14.json 2> gpurun_out/bench_c14.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_c14.json').read().strip().splitlines()[-1])
    print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'fmha frac', round(d['roofline']['frac'],3))
    for k,v in d['kernels'].items(): print(' ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
except Exception as e: print('parse failed', e); print(open('gpurun_out/bench_c14.err').read()[-2000:])
PY
