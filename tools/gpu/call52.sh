#!/bin/bash
# round 2, call 52: evidence of the final tree: plain bench, then the ncu launch list of the same command and one full capture
# of the attention forward
mkdir -p gpurun_out
timeout 900 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_c52.json 2> gpurun_out/bench_c52.err; echo "bench exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_final2.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"fmha_fwd_kernel" -s 2 -c 1 -o gpurun_out/r02_fmha_final python tools/kernel_bench.py fmha --iters 1 > gpurun_out/ncu_fmha_final.log 2>&1; echo "ncu exit $?"
