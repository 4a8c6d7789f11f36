#!/usr/bin/env python
"""Host -> device bandwidth of one 679 MB step (the bench's fp32 feature batch) from (a) torch pinned memory, (b) pinned memory
allocated write-combined (cudaHostAllocWriteCombined: no CPU cache snooping on the PCIe reads), (c) several chunks on two
streams.  CUDA events around 10 copies each."""
import ctypes
import json

import numpy as np
import torch

N = 678_732_064  # bytes per step (bench.py e2e.h2d_bytes_per_step)
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
dst = torch.empty(N, dtype=torch.uint8, device=dev)
out = {}


def time_copies(src, n=10, streams=1, chunks=1):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    step = (N + chunks - 1) // chunks
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        for c in range(chunks):
            lo, hi = c * step, min(N, (c + 1) * step)
            s = ss[c % streams]
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                dst[lo:hi].copy_(src[lo:hi], non_blocking=True)
        for s in ss:
            torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return N * n / (e0.elapsed_time(e1) * 1e-3) / 1e9


pinned = torch.empty(N, dtype=torch.uint8).pin_memory()
pinned.fill_(1)
out["torch_pinned_gbs"] = time_copies(pinned)
out["torch_pinned_4chunks_2streams_gbs"] = time_copies(pinned, streams=2, chunks=4)
try:
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(N), ctypes.c_uint(0x04))  # cudaHostAllocWriteCombined
    assert rc == 0, rc
    arr = np.ctypeslib.as_array((ctypes.c_uint8 * N).from_address(p.value))
    arr[:] = 1
    wc = torch.from_numpy(arr)
    out["write_combined_gbs"] = time_copies(wc)
    out["write_combined_is_pinned_for_torch"] = bool(wc.is_pinned())
    rt.cudaFreeHost(p)
except Exception as e:  # noqa: BLE001
    out["write_combined_error"] = repr(e)[:200]
print(json.dumps(out))
