"""Evaluation metrics on the device: `atiou` restates `calculate_tiou` (reference utils/metrics.py:82-111)
and the averaging loop of inference.py:45-55 over the gathered fixed-slot segment lists (see
repurpose_b200/scheduler.py for the slot layout), in float64 with the reference's operation order, so
a sharded 10K-video evaluation ends with one tiny device->host copy."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, cur_stream, ptr

THRESHOLDS = (0.5, 0.6, 0.7, 0.8, 0.9)  # inference.py:46


def atiou(slots: torch.Tensor, gt_segments, thresholds=THRESHOLDS):
    """slots [n, 1+4K] f32 CUDA; gt_segments: list (one per video) of [[start, end], ...] Python floats.
    Returns (average tIoU, {threshold: mean precision}, per-video precision tensor [n, n_thr])."""
    if not slots.is_cuda:
        raise _lib.RepurposeError("atiou needs the gathered slots on a CUDA device (no CPU path)")
    dev = slots.device
    n = slots.shape[0]
    K = (slots.shape[1] - 1) // 4
    gmax = max(1, max((len(g) for g in gt_segments), default=1))
    gt = torch.zeros(n, gmax, 2, dtype=torch.float64)
    cnt = torch.zeros(n, dtype=torch.int32)
    for i, g in enumerate(gt_segments):
        cnt[i] = len(g)
        if len(g):
            gt[i, :len(g)] = torch.tensor(g, dtype=torch.float64)
    gt, cnt = gt.to(dev), cnt.to(dev)
    thr = torch.tensor(list(thresholds), dtype=torch.float64, device=dev)
    per_video = torch.empty(n, len(thresholds), dtype=torch.float64, device=dev)
    out = torch.empty(len(thresholds) + 1, dtype=torch.float64, device=dev)
    slots = slots.contiguous().float()
    with torch.cuda.device(dev):
        check(_lib.load().rp_atiou(ptr(slots), n, K, ptr(gt), ptr(cnt), gmax, ptr(thr), len(thresholds),
                                   ptr(per_video), ptr(out), cur_stream()), "rp_atiou")
    o = out.tolist()
    return o[-1], dict(zip(thresholds, o[:-1])), per_video
