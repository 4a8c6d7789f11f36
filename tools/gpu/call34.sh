#!/bin/bash
# round 2, call 34: (a) one launch sequence over the batch vs two / four half batches on separate streams,
# (b) training-step comparator: torch eager (fp32 / bf16 autocast) + cuDNN SDPA fwd/bwd vs TrainStep
mkdir -p gpurun_out
for p in 2 4; do timeout 600 python tools/two_stream_probe.py --parts $p 2>&1 | tail -3; done
timeout 900 python tools/train_comparator.py --out gpurun_out/train_comparator.json 2>&1 | tail -5
