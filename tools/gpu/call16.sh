#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train_step.py tests/test_train_pieces.py -q -x -p no:cacheprovider 2>&1 | tail -5
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -k "head_out or layernorm" -p no:cacheprovider 2>&1 | tail -2
timeout 600 python tools/train_bench.py --B 16 --T 1801 --steps 3 > gpurun_out/train_bench_b16.json 2> gpurun_out/train_bench_b16.err; echo "train bench exit $?"; cat gpurun_out/train_bench_b16.json; tail -3 gpurun_out/train_bench_b16.err
timeout 600 python tools/train_bench.py --B 32 --T 1801 --steps 3 > gpurun_out/train_bench_b32.json 2> gpurun_out/train_bench_b32.err; echo "train bench exit $?"; cat gpurun_out/train_bench_b32.json; tail -3 gpurun_out/train_bench_b32.err
