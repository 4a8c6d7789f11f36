"""ORACLE (test infrastructure, see oracle/__init__.py): NumPy restatement of the reference's
Gaussian Soft-NMS over 1-D intervals, `soft_nms_intervals_cpu` (models/softnms.py:3-38), written
from the semantics in SURVEY.md Appendix B rather than from the reference text:

  * candidates live in three parallel arrays (begin, end, original index) that are permuted
    together; `len0[pos] = end - begin` is computed once and is NOT permuted (models/softnms.py:13
    computes `lengths` once, the swap at :24 only touches the offsets rows) — "stale lengths";
  * round i reads `tscore = scores[i]` BEFORE the swap (:18) and that value drives the selected
    counter (:26-29); the break happens before the decay of that round;
  * arg-max over the tail takes the first index on ties (np.argmax, :22); swap only if strictly
    smaller (:23);
  * decay: ov = max(min(e_i, e_j) - max(b_i, b_j), 0); total = (len0[i] + len0[j]) - ov;
    w = exp(-(ov/total)^2 / sigma) in float32 (:30-36);
  * keep = original indices of rows whose score > thresh, in the final (permuted) order, first
    min(max_seg_num, N) of them (:37).
The inputs are never modified (the reference aliases CPU tensors).
"""
from __future__ import annotations

import numpy as np


def soft_nms_intervals_oracle(scores, segments, sigma=0.5, thresh=0.001, max_seg_num=20,
                              return_scores=False):
    sc = np.array(scores, dtype=np.float32, copy=True).reshape(-1)
    seg = np.array(segments, dtype=np.float32, copy=True).reshape(-1, 2)
    n = sc.shape[0]
    begin = seg[:, 0].copy()
    end = seg[:, 1].copy()
    orig = np.arange(n, dtype=np.int64)
    len0 = end - begin                       # indexed by position, never permuted
    sig = np.float32(sigma)
    thr = np.float32(thresh)
    m = min(int(max_seg_num), n)
    cnt = 0
    for i in range(n):
        tscore = sc[i]                       # pre-swap value
        if i != n - 1:
            j = i + 1 + int(np.argmax(sc[i + 1:]))
            if tscore < sc[j]:
                begin[i], begin[j] = begin[j], begin[i]
                end[i], end[j] = end[j], end[i]
                orig[i], orig[j] = orig[j], orig[i]
                sc[i], sc[j] = sc[j], sc[i]
        if tscore > thr:
            cnt += 1
            if cnt >= m:
                break
        ov = np.maximum(np.minimum(end[i], end[i + 1:]) - np.maximum(begin[i], begin[i + 1:]),
                        np.float32(0))
        total = (len0[i] + len0[i + 1:]) - ov
        with np.errstate(divide="ignore", invalid="ignore"):
            r = ov / total
            w = np.exp(-(r * r) / sig)
        sc[i + 1:] = w * sc[i + 1:]
    sel = sc > thr
    keep = orig[sel][:m]
    if return_scores:
        return keep, sc[sel][:m].copy()
    return keep
