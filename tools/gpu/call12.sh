#!/bin/bash
# round 2, call 12: training step end to end (gradient parity), first timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_step.py -q -x -s -p no:cacheprovider > gpurun_out/test_train_step.log 2>&1; echo "train step exit $?"; tail -n 40 gpurun_out/test_train_step.log
timeout 600 python tools/train_bench.py --B 8 --T 1801 --steps 3 > gpurun_out/train_bench_b8.json 2> gpurun_out/train_bench_b8.err; echo "train bench exit $?"; cat gpurun_out/train_bench_b8.json; tail -5 gpurun_out/train_bench_b8.err
