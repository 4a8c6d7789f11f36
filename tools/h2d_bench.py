"""Micro-benchmark: pinned host -> device bandwidth: default pinned vs write-combined pinned memory,
idle and while a GEMM keeps the GPU busy."""
import ctypes, time
import numpy as np, torch
n = 679 * 1024 * 1024 // 4
rt = ctypes.CDLL("libcudart.so.12")
def wc_tensor(n_float):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n_float * 4), ctypes.c_uint(4)) == 0   # cudaHostAllocWriteCombined
    arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(n_float,))
    return torch.from_numpy(arr)
dst = torch.empty(n, dtype=torch.float32, device="cuda")
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
copy_stream = torch.cuda.Stream()
for name, src in (("default pinned", torch.empty(n, dtype=torch.float32).pin_memory()), ("write-combined", wc_tensor(n))):
    src[:1024] = 1.0
    for busy in (False, True):
        def run():
            with torch.cuda.stream(copy_stream):
                dst.copy_(src, non_blocking=True)
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if busy:
            for _ in range(40):
                a @ a
        with torch.cuda.stream(copy_stream):
            e0.record(copy_stream)
        for _ in range(5):
            run()
        with torch.cuda.stream(copy_stream):
            e1.record(copy_stream)
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) / 5 / 1e3
        print(f"{name:16s} busy={busy}: {n * 4 / dt / 1e9:.1f} GB/s ({dt * 1e3:.2f} ms per 679 MB)")
