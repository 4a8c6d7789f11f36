#!/bin/bash
# round 2, call 53: attention forward: partial query tiles at the end of the launch (shipped) vs in order: time alone, in the
# step, and DRAM traffic (ncu)
mkdir -p gpurun_out
for rep in 1 2; do for v in main inorder; do
  if [ $v = main ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_$v.so; fi
  for T in 1801; do echo -n "$v T=$T: "; timeout 300 python tools/kernel_bench.py fmha --iters 30 --T $T 2>&1 | tr -d '\n '; echo; done
done; done
for v in main inorder main inorder; do
  if [ $v = main ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_$v.so; fi
  timeout 900 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'fmha ms', round(d['kernels']['fmha']['ms_per_step'],3), d['clocks']['sm_mhz'])"
done
export RP_LIB_PATH=ab/lib_inorder.so
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"fmha_fwd_kernel" -s 2 -c 1 python tools/kernel_bench.py fmha --iters 1 2>&1 | grep -E "dram__|gpu__time"
