#!/bin/bash
# round 2, call 10: GPU tests, residual-slab depth A/B on the LayerNorm-fused GEMMs, row-wise kernels, bench, ncu DRAM captures
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh 2>&1 | grep -E "^== |passed|failed|error" | head -40
for s in 4 6 8; do
  echo "== RP_RESID_SLABS=$s"
  RP_RESID_SLABS=$s timeout 300 python tools/kernel_bench.py gemmln --iters 20 2>&1 | tr -d '\n '; echo
done
echo "== resid (unfused) K=512 slabs 4 vs 6"
for s in 4 6; do RP_RESID_SLABS=$s timeout 300 python tools/kernel_bench.py gemm --iters 10 2>&1 | tr -d '\n ' | grep -o '"gemm_out(epi3)[^}]*}\|"gemm_ff2(epi3)[^}]*}'; echo; done
echo "== rowwise"
timeout 300 python tools/kernel_bench.py rowwise ln --iters 20 2>&1 | tr -d '\n '; echo
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 3 --no-extras > gpurun_out/bench_c10.json 2> gpurun_out/bench_c10.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_c10.json').read().strip().splitlines()[-1])
    print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'fmha frac', round(d['roofline']['frac'],3), 'cpu', d['cpu_baseline'])
    for k,v in d['kernels'].items(): print(' ', k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
except Exception as e: print('parse failed', e); print(open('gpurun_out/bench_c10.err').read()[-2000:])
PY
echo "== ncu rowwise"
timeout 600 ncu --set full --clock-control none -k regex:"layernorm512|concat_cast|head_out" -c 12 -o gpurun_out/r02_rowwise python tools/kernel_bench.py rowwise --iters 1 > gpurun_out/ncu_rowwise.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --set full --clock-control none -k regex:"decode_nms" -c 2 -o gpurun_out/r02_decode python __graft_entry__.py smoke > gpurun_out/ncu_decode.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/*.ncu-rep
