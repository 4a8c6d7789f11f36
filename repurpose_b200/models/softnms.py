"""Gaussian Soft-NMS over 1-D intervals on the GPU, behind the reference's call signature.

Replaces `soft_nms_intervals_cpu` (reference models/softnms.py:3-38).  The reference moves the
candidates to the host and loops in NumPy; here one CTA per video runs the identical round
structure (arg-max, swap, early stop at `max_seg_num`, IoU decay, final `score > thresh` filter,
including the stale-`lengths` and pre-swap-`tscore` quirks) with the
candidates in shared memory.  Unlike the CPU-tensor case of the reference, the caller's tensors
are never mutated (the reference's own behaviour for CUDA tensors).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .._lib import check, cur_stream, ptr

MAX_CANDIDATES = 8192


def soft_nms_batched(scores, segments, n, max_seg, sigma, thresh, kcap=None):
    """scores [B,Nmax] f32, segments [B,Nmax,2] f32, n [B] i32, max_seg [B] i32 (all CUDA) ->
    (keep [B,Kcap] i32, decayed scores [B,Kcap] f32, counts [B] i32), device tensors."""
    dev = scores.device
    if dev.type != "cuda":
        raise _lib.RepurposeError("soft_nms_batched needs CUDA tensors (no CPU path)")
    B, nmax = scores.shape
    if kcap is None:
        kcap = max(1, int(max_seg.max().item()))
    keep = torch.full((B, kcap), -1, dtype=torch.int32, device=dev)
    kscores = torch.zeros(B, kcap, dtype=torch.float32, device=dev)
    counts = torch.zeros(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().rp_soft_nms(ptr(scores), ptr(segments), ptr(n), ptr(max_seg), B, nmax,
                                      float(sigma), float(thresh), kcap, ptr(keep), ptr(kscores),
                                      ptr(counts), cur_stream()), "rp_soft_nms")
    return keep, kscores, counts


def _single(out_cls_logits, out_offsets, sigma, thresh, max_seg_num):
    if not torch.cuda.is_available():
        raise _lib.RepurposeError("Soft-NMS runs on the GPU only; no CUDA device is available")
    dev = out_cls_logits.device if out_cls_logits.is_cuda else torch.device("cuda")
    n = int(out_offsets.shape[0])
    if n == 0 or max_seg_num <= 0:
        return None, dev
    if n > MAX_CANDIDATES:
        raise _lib.RepurposeError(f"{n} candidates exceed the kernel limit of {MAX_CANDIDATES}")
    scores = out_cls_logits.detach().to(device=dev, dtype=torch.float32).reshape(1, n).contiguous()
    segs = out_offsets.detach().to(device=dev, dtype=torch.float32).reshape(1, n, 2).contiguous()
    n_t = torch.tensor([n], dtype=torch.int32, device=dev)
    m_t = torch.tensor([int(max_seg_num)], dtype=torch.int32, device=dev)
    keep, kscores, counts = soft_nms_batched(scores, segs, n_t, m_t, sigma, thresh,
                                             kcap=max(1, min(int(max_seg_num), n)))
    k = int(counts[0].item())
    return (keep[0, :k], kscores[0, :k]), dev


def soft_nms_intervals(out_cls_logits, out_offsets, sigma=0.5, thresh=0.001, max_seg_num=20):
    """Device-resident variant (the name the reference mentions in a commented-out line,
    models/MMCTransformer.py:264): returns a CUDA int64 index tensor."""
    r, dev = _single(out_cls_logits, out_offsets, sigma, thresh, max_seg_num)
    if r is None:
        return torch.empty(0, dtype=torch.int64, device=dev)
    return r[0].to(torch.int64)


def soft_nms_intervals_cpu(out_cls_logits, out_offsets, sigma=0.5, thresh=0.001, max_seg_num=20):
    """Same signature and return type as the reference (np.ndarray[int64] of kept original
    indices in selection order); the work is done by the CUDA kernel."""
    r, _ = _single(out_cls_logits, out_offsets, sigma, thresh, max_seg_num)
    if r is None:
        return np.empty(0, dtype=np.int64)
    return r[0].cpu().numpy().astype(np.int64)
