"""numpy restatement of the dropout mask function of csrc/dropout.cuh (test helper, not product code):

    pair k = element >> 1;  h = fmix32(k * 0x9E3779B1 + key_a) ^ key_b
    element 2k kept iff (h & 0xffff) >= thr, element 2k+1 iff (h >> 16) >= thr, thr = round(p * 65536)
"""
import numpy as np


def fmix32(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x85EBCA6B)
    x ^= x >> np.uint32(13)
    x *= np.uint32(0xC2B2AE35)
    x ^= x >> np.uint32(16)
    return x


def keep_mask(key_a, key_b, p, n):
    """uint8 [n]: 1 = kept"""
    pairs = (n + 1) // 2
    with np.errstate(over="ignore"):
        k = np.arange(pairs, dtype=np.uint32)
        h = fmix32(k * np.uint32(0x9E3779B1) + np.uint32(key_a)) ^ np.uint32(key_b)
    thr = np.uint32(min(65535, max(1, int(p * 65536.0 + 0.5))))
    out = np.empty(2 * pairs, dtype=np.uint8)
    out[0::2] = (h & np.uint32(0xFFFF)) >= thr
    out[1::2] = (h >> np.uint32(16)) >= thr
    return out[:n]
