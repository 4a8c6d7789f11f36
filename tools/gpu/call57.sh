#!/bin/bash
# round 2, call 57: DRAM bytes per kernel over bench steps (cold-cache replays): is any kernel far above its algorithmic bytes?
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 130 -c 125 --csv --log-file gpurun_out/r02_dram_per_kernel.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1; echo "ncu exit $?"
