#!/usr/bin/env python
"""Write a small synthetic dataset in the reference's on-disk format (label JSON + three .npy feature files
per video + a Repurpose-style YAML config + a random-init checkpoint), e.g. to try `repurpose_b200.infer`:

    python tools/make_synth_dataset.py /tmp/synth --videos 24
    python -m repurpose_b200.infer --config_path /tmp/synth/cfg.yaml --resume /tmp/synth/ckpt.pth
    torchrun --nproc-per-node 2 -m repurpose_b200.infer --config_path /tmp/synth/cfg.yaml --resume /tmp/synth/ckpt.pth
"""
import argparse, json, sys
from pathlib import Path

import numpy as np
import torch
import yaml

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from repurpose_b200 import synth  # noqa: E402
from repurpose_b200.models.MMCTransformer import MMCTransformer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("out")
ap.add_argument("--videos", type=int, default=24)
ap.add_argument("--layers", type=int, default=2)
ap.add_argument("--max-len", type=int, default=600)
a = ap.parse_args()
out = Path(a.out)
rng = np.random.default_rng(0)
dims = {"video_path": 512, "audio_path": 2048, "text_path": 384}
dirs = {k: out / k for k in dims}
for d in dirs.values():
    d.mkdir(parents=True, exist_ok=True)
labels = []
for i, n in enumerate(synth.sample_lengths(a.videos, seed=3, t_max=a.max_len)):
    n = max(n, 70)
    for k, d in dirs.items():
        np.save(d / f"vid{i}.npy", rng.normal(size=(n, dims[k])).astype(np.float32))
    gt = sorted(sorted(rng.uniform(0, n, size=2).tolist()) for _ in range(int(rng.integers(1, 6))))
    labels.append({"youtube_id": f"vid{i}", "timeRange": [0, float(n - 1)], "timeRangeOffset": [0, float(n - 1)],
                   "segments": gt, "segmentsOffset": gt})
(out / "test.json").write_text(json.dumps(labels))
model_cfg = dict(synth.MODEL_CFG, self_num_layers=a.layers)
cfg = {"test_dataset": {"label_path": str(out / "test.json"), **{k: str(v) for k, v in dirs.items()}},
       "model": model_cfg, "test_cfg": dict(synth.TEST_CFG)}
(out / "cfg.yaml").write_text(yaml.safe_dump(cfg))
torch.manual_seed(0)
m = MMCTransformer(**model_cfg)
torch.save({"model": synth.bias_reg_head({k: v.clone() for k, v in m.state_dict().items()})}, out / "ckpt.pth")
print(f"wrote {a.videos} videos, cfg.yaml and ckpt.pth under {out}")
