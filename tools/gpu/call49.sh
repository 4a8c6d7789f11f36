#!/bin/bash
# round 2, call 49 (2 GPUs): bucketed, overlapped gradient all-reduce against the flat one
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/allreduce_overlap_check.py 2>&1 | tail -5
