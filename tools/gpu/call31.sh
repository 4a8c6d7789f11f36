#!/bin/bash
# round 2, call 31: persistent attention with lean waits in the auxiliary warps
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "fmha" 2>&1 | tail -2
echo "== shipped"; for T in 1801 1792 8192; do B=32; if [ $T = 8192 ]; then B=4; fi; RP_FMHA_PERSIST=0 timeout 300 python tools/kernel_bench.py fmha --iters 20 --T $T --B $B 2>&1 | tr -d '\n '; echo " T=$T"; done
for sg in 0 2600; do echo "== persist stagger $sg"; for T in 1801 1792 8192; do B=32; if [ $T = 8192 ]; then B=4; fi; RP_FMHA_STAGGER=$sg RP_FMHA_PERSIST=1 timeout 300 python tools/kernel_bench.py fmha --iters 20 --T $T --B $B 2>&1 | tr -d '\n '; echo " T=$T"; done; done
