#!/bin/bash
# round 2, call 1: ping-pong FMHA correctness + A/B timing, SDPA comparator, a first bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
t() { local name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "== $name: exit $?"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
t fmha_pp_tests python -m pytest tests/test_gpu_kernels.py -q -m gpu -k fmha -p no:cacheprovider
RP_FMHA_PP=0 t fmha_pp0_tests python -m pytest tests/test_gpu_kernels.py -q -m gpu -k fmha -p no:cacheprovider
RP_FMHA_EMU=2 t fmha_emu2_tests python -m pytest tests/test_gpu_kernels.py -q -m gpu -k fmha_key -p no:cacheprovider
for v in "RP_FMHA_V=1" "RP_FMHA_V=2 RP_FMHA_PP=0" "RP_FMHA_V=2" "RP_FMHA_V=2 RP_FMHA_EMU=0" "RP_FMHA_V=2 RP_FMHA_EMU=2"; do
  echo "---- $v"; env $v timeout 300 python tools/kernel_bench.py fmha --iters 20 2>&1 | tail -n 6
done | tee gpurun_out/fmha_ab.log
TAILN=40 t comparator_sdpa python tools/comparator.py --skip-model --iters 10
TAILN=5 t bench python bench.py
RP_FMHA_V=1 TAILN=5 t bench_v1 python bench.py
