#!/bin/bash
mkdir -p gpurun_out
t() { local name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "== $name: exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
t fmha_v1tail_tests python -m pytest tests/test_gpu_kernels.py -q -m gpu -k fmha -p no:cacheprovider
RP_FMHA_NQ=2 t fmha_v1tail_nq2_tests python -m pytest tests/test_gpu_kernels.py -q -m gpu -k fmha -p no:cacheprovider
for v in "RP_FMHA_V=1" "RP_FMHA_V=3" "RP_FMHA_V=1 RP_FMHA_EMU=0" "RP_FMHA_V=1 RP_FMHA_EMU=2" "RP_FMHA_V=1"; do
  echo "---- $v"; env $v timeout 300 python tools/kernel_bench.py fmha --iters 20 2>&1 | tail -n 6
done | tee gpurun_out/fmha_ab6.log
echo "---- T=1792 (no tails)"; python tools/kernel_bench.py fmha --iters 20 --T 1792 | tail -n 6
TAILN=3 t bench_v1tail python bench.py
TAILN=8 t model_tests python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -x
