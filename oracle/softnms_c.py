"""ORACLE (test infrastructure): ctypes loader for the plain-C Soft-NMS restatement
(oracle/softnms_c.c, built by `make -C oracle` / __graft_entry__.build())."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "_build" / "liboracle_softnms.so"
_lib = None


def build() -> Path:
    subprocess.run(["make", "-C", str(_DIR), "-s"], check=True)
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not _SO.exists():
            build()
        _lib = C.CDLL(str(_SO))
        _lib.rp_oracle_soft_nms.restype = C.c_int
        _lib.rp_oracle_soft_nms.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                            C.c_int, C.c_void_p, C.c_void_p]
    return _lib


def soft_nms_intervals_c(scores, segments, sigma=0.5, thresh=0.001, max_seg_num=20,
                         return_scores=False):
    sc = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1)
    seg = np.ascontiguousarray(segments, dtype=np.float32).reshape(-1, 2)
    n = sc.shape[0]
    cap = max(1, min(int(max_seg_num), n))
    keep = np.zeros(cap, dtype=np.int64)
    ks = np.zeros(cap, dtype=np.float32)
    k = _load().rp_oracle_soft_nms(sc.ctypes.data, seg.ctypes.data, n, float(sigma), float(thresh),
                                   int(max_seg_num), keep.ctypes.data, ks.ctypes.data)
    if return_scores:
        return keep[:k].copy(), ks[:k].copy()
    return keep[:k].copy()
