#!/bin/bash
mkdir -p gpurun_out
export RP_LIB_PATH=ab/lib_trace.so
echo "==== PP=1"; timeout 300 python tools/fmha_pp_trace.py 2>&1 | tail -n 50
echo "==== PP=0"; RP_FMHA_PP=0 timeout 300 python tools/fmha_pp_trace.py 2>&1 | tail -n 50
echo "==== EMU=0"; RP_FMHA_EMU=0 timeout 300 python tools/fmha_pp_trace.py 2>&1 | tail -n 50
