#!/bin/bash
mkdir -p gpurun_out
bash tools/run_gpu_tests.sh
python tools/kernel_bench.py fmha --iters 20 | tail -n 6
timeout 600 python bench.py > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; echo "bench exit $?"; tail -c 600 gpurun_out/bench_r02_a.json
timeout 900 python tools/comparator.py --iters 3 > gpurun_out/comparator_full.log 2>&1; echo "comparator exit $?"; tail -n 30 gpurun_out/comparator_full.log
