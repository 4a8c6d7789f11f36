#!/bin/bash
# round 2, call 41: probe: dQ MMA with M = 64, accumulator rows assumed in TMEM lanes 0-63 (drain warps 8, 9)
export RP_LIB_PATH=ab/lib_dqm64.so
timeout 600 python -m pytest tests/test_gpu_train_kernels.py -q -m gpu -p no:cacheprovider -k "fmha_backward" 2>&1 | tail -6
for B in 16 32; do echo -n "dqm64 B=$B: "; timeout 300 python tools/kernel_bench.py fmhabwd --iters 20 --B $B 2>&1 | tr -d '\n '; echo; done
unset RP_LIB_PATH
for B in 16 32; do echo -n "main B=$B: "; timeout 300 python tools/kernel_bench.py fmhabwd --iters 20 --B $B 2>&1 | tr -d '\n '; echo; done
