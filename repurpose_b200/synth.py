"""Seeded synthetic inputs / weights shared by the golden generator, the tests and bench.py.

Shapes and thresholds come from the reference's configs/Repurpose.yaml (:22-32 model, :52-61
test_cfg); the batch dict layout is what dataset/RepurposeClip.py:preprocessing()/collate_fn_test
hand to the model.
"""
from __future__ import annotations

import numpy as np
import torch

MODEL_CFG = dict(vis_dim=512, aud_dim=2048, text_dim=384, d_model=512, self_num_layers=16,
                 text_num_layers=3, cross_num_layers=3, num_heads=8)
TEST_CFG = dict(pre_nms_topk=1000, pre_nms_thresh=0.5, duration_thresh=10, duration_thresh_max=90,
                max_seg_per_min=0.3, nms_sigma=0.5, min_score=0.01)
MAX_SEQ_LEN = 1801  # dataset maximum under Repurpose.yaml's label files

# deciles (0,5,10,...,100 %) of data/test.json timeRangeOffset lengths (min 28, p50 1304, max 1801)
_LEN_Q = np.array([0, 5, 10, 20, 30, 40, 50, 60, 70, 80, 90, 95, 100], dtype=np.float64)
_LEN_V = np.array([28, 257, 380, 590, 767, 989, 1304, 1738, 1801, 1801, 1801, 1801, 1801],
                  dtype=np.float64)


def sample_lengths(n: int, seed: int, t_max: int = MAX_SEQ_LEN) -> list[int]:
    """Video lengths drawn from the test-split length distribution (piecewise-linear quantiles)."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(0, 100, size=n)
    v = np.interp(q, _LEN_Q, _LEN_V) * (t_max / float(MAX_SEQ_LEN))
    return [int(max(1, min(t_max, round(x)))) for x in v]


def make_batch(lens, seed: int, T: int | None = None, cfg=MODEL_CFG, pin: bool = False, smooth: int = 0):
    """Collated batch on the host: feats ~ N(0,1) fp32 (zero in padded steps like preprocessing()),
    masks [B,1,T] bool left-aligned, labels/segments zeros, video_id, duration = lens.
    smooth = W > 1: the features are a W-step moving average of white noise (unit variance kept), i.e.
    temporally correlated like real 1-fps CLIP / PANNs / MiniLM features, instead of independent per step."""
    lens = [int(x) for x in lens]
    B = len(lens)
    T = int(T if T is not None else max(lens))
    g = torch.Generator().manual_seed(int(seed))
    valid = (torch.arange(T)[None, :] < torch.tensor(lens)[:, None])
    batch = {}
    for key, dim in (("visual_feats", cfg["vis_dim"]), ("audio_feats", cfg["aud_dim"]),
                     ("text_feats", cfg["text_dim"])):
        x = torch.randn(B, T, dim, generator=g, dtype=torch.float32)
        if smooth > 1:
            w = int(smooth) | 1
            x = torch.nn.functional.avg_pool1d(x.transpose(1, 2), w, 1, w // 2, count_include_pad=True)
            x = (x.transpose(1, 2) * (w ** 0.5)).contiguous()
        x = x * valid[:, :, None]
        batch[key] = x.pin_memory() if pin else x
    batch["masks"] = valid[:, None, :].clone()
    batch["labels"] = torch.zeros(B, T)
    batch["segments"] = torch.zeros(B, T, 2)
    batch["video_id"] = [f"vid{seed}_{i}" for i in range(B)]
    batch["duration"] = lens
    return batch


def bias_reg_head(state_dict):
    """Random-init reg_head emits offsets ~0-2, so every candidate fails duration_thresh > 10 and
    Soft-NMS sees N = 0.  Same remedy as the survey: scale/bias the last layer so
    durations spread over (10, 90)."""
    sd = dict(state_dict)
    sd["reg_head.7.weight"] = state_dict["reg_head.7.weight"] * 40.0
    sd["reg_head.7.bias"] = torch.full_like(state_dict["reg_head.7.bias"], 20.0)
    return sd


def make_candidates(n: int, T: int, seed: int):
    """Soft-NMS candidates: centres U(0,T), lengths U(10.5,89.5), scores U(0.5,1)
    sorted descending.  Returns (scores [n] f32, segments [n,2] f32) numpy arrays."""
    rng = np.random.default_rng(seed)
    c = rng.uniform(0, T, size=n)
    ln = rng.uniform(10.5, 89.5, size=n)
    segs = np.stack([c - ln / 2, c + ln / 2], axis=1).astype(np.float32)
    scores = np.sort(rng.uniform(0.5, 1.0, size=n).astype(np.float32))[::-1].copy()
    return scores, segs


def make_gt_segments(T: int, seed: int):
    """Ground truth for AtIoU scoring: 1..34 random segments, median length ~36 s."""
    rng = np.random.default_rng(seed)
    k = int(rng.integers(1, 35))
    starts = rng.uniform(0, max(1.0, T - 40.0), size=k)
    lens = np.clip(rng.lognormal(mean=np.log(36.0), sigma=0.5, size=k), 8.0, 120.0)
    return [[float(s), float(min(T, s + l))] for s, l in zip(starts, lens)]


def max_seg_num(duration: int, max_seg_per_min: float) -> int:
    """models/MMCTransformer.py:255-257"""
    return int(np.ceil((int(duration) // 60) * max_seg_per_min))
