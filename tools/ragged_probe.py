#!/usr/bin/env python
"""Per-kernel view of a caller-built ragged batch (bench.py extra `ragged_unsorted_batch`): run it once with and once without
the padding skip; meant to be run under `ncu --metrics gpu__time_duration.sum` (tools/summarize_ncu.py launches ...).

    python tools/ragged_probe.py [--skip 0|1]
"""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from repurpose_b200 import synth  # noqa: E402
from repurpose_b200.models.MMCTransformer import MMCTransformer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--skip", type=int, default=1)
ap.add_argument("--iters", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = MMCTransformer(**synth.MODEL_CFG)
model.load_state_dict(synth.bias_reg_head({k: v.clone() for k, v in model.state_dict().items()}))
model = model.to(dev).eval()
lens = synth.sample_lengths(32, seed=4242)
host = synth.make_batch(lens, seed=4243, T=1801)
devb = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
model.set_skip_padding(bool(a.skip))
for _ in range(1 + a.iters):
    model.inference_device(devb, synth.TEST_CFG)
torch.cuda.synchronize()
