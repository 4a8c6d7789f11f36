#!/bin/bash
# round 2, call 30: persistent attention: start stagger sweep
mkdir -p gpurun_out
echo "== shipped"; for T in 1801 8192; do B=32; if [ $T = 8192 ]; then B=4; fi; RP_FMHA_PERSIST=0 timeout 300 python tools/kernel_bench.py fmha --iters 20 --T $T --B $B 2>&1 | tr -d '\n '; echo " T=$T"; done
for sg in 0 1300 2600 6000 20000; do echo "== persist stagger $sg"; for T in 1801 8192; do B=32; if [ $T = 8192 ]; then B=4; fi; RP_FMHA_STAGGER=$sg RP_FMHA_PERSIST=1 timeout 300 python tools/kernel_bench.py fmha --iters 20 --T $T --B $B 2>&1 | tr -d '\n '; echo " T=$T"; done; done
