#!/bin/bash
# round 2, call 47: attention CTAs laid out longest video first for ragged batches: model tests + the ragged batch timing
timeout 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -3
python - <<'PY'
import torch, json, sys
sys.path.insert(0, '.')
import bench
from repurpose_b200 import synth
from repurpose_b200.models.MMCTransformer import MMCTransformer
dev = torch.device('cuda', 0)
torch.manual_seed(0)
model = MMCTransformer(**synth.MODEL_CFG)
model.load_state_dict(synth.bias_reg_head({k: v.clone() for k, v in model.state_dict().items()}))
model = model.to(dev).eval()
for rep in range(2):
    print(json.dumps(bench.ragged_unsorted(model, dev, iters=10)))
PY
