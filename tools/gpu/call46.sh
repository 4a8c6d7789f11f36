#!/bin/bash
# round 2, call 46: per-kernel times of the caller-built ragged batch with / without the padding skip (ncu launch lists)
mkdir -p gpurun_out
for s in 0 1; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ragged_launches_skip$s.csv python tools/ragged_probe.py --skip $s > /dev/null 2>&1; echo "ncu exit $?"
done
