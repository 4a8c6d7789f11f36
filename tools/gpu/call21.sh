#!/bin/bash
# round 2, call 21 (2 GPUs): traceback of the two-device test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -p no:cacheprovider -k "two_devices" --tb=long 2>&1 | tail -80 > gpurun_out/two_dev.log
tail -60 gpurun_out/two_dev.log
