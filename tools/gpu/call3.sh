#!/bin/bash
mkdir -p gpurun_out
t() { local name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "== $name: exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
export RP_FMHA_V=3
t fmha_k64_tests python -m pytest tests/test_gpu_kernels.py -q -m gpu -k fmha -p no:cacheprovider
for v in "RP_FMHA_V=1" "RP_FMHA_V=3" "RP_FMHA_V=3 RP_FMHA_EMU=0" "RP_FMHA_V=3 RP_FMHA_EMU=2" "RP_FMHA_V=1"; do
  echo "---- $v"; env $v timeout 300 python tools/kernel_bench.py fmha --iters 20 2>&1 | tail -n 6
done | tee gpurun_out/fmha_ab3.log
TAILN=3 t bench_v3 python bench.py
