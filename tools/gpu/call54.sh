#!/bin/bash
# round 2, call 54: attention forward: how many (batch, head) pairs' partial query tiles to defer to the end of the launch
# (0 = in order, 48, 96 = main build, 256 = all, the previous layout): time alone, in the step, DRAM traffic; tests on main
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k fmha 2>&1 | tail -2
for rep in 1 2; do for v in d0 d48 main d256; do
  if [ $v = main ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_$v.so; fi
  echo -n "$v: "; timeout 300 python tools/kernel_bench.py fmha --iters 30 2>&1 | tr -d '\n '; echo
done; done
for rep in 1 2; do for v in d0 d48 main d256; do
  if [ $v = main ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_$v.so; fi
  timeout 900 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'fmha ms', round(d['kernels']['fmha']['ms_per_step'],3), d['clocks']['sm_mhz'])"
done; done
for v in d48 main; do
  if [ $v = main ]; then unset RP_LIB_PATH; else export RP_LIB_PATH=ab/lib_$v.so; fi
  echo "ncu $v"; timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"fmha_fwd_kernel" -s 2 -c 1 python tools/kernel_bench.py fmha --iters 1 2>&1 | grep -E "dram__|gpu__time"
done
