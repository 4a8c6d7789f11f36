#!/usr/bin/env python
"""Micro-benchmark of the individual kernels at the bench.py shapes (B=32, T=1801), timed with CUDA
events.  Used while optimising and as the target command for `ncu -k regex:<kernel>`.

    python tools/kernel_bench.py [fmha] [gemm] [ln] [--iters N] [--B 32] [--T 1801]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from repurpose_b200 import _lib  # noqa: E402
from repurpose_b200._lib import check, cur_stream, ptr  # noqa: E402

LOG2E = 1.4426950408889634


def timeit(fn, iters, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="*", default=["fmha", "gemm", "ln"])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--T", type=int, default=1801)
    ap.add_argument("--lens", default=None, help="comma-separated kv lens (default all T)")
    args = ap.parse_args()
    lib = _lib.load()
    dev = "cuda"
    B, T, H, D = args.B, args.T, 8, 512
    M = B * T
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    out = {}
    if "fmha" in args.what:
        qkv = torch.randn(B, T, 3 * D, device=dev)
        qkv[..., :D] *= LOG2E / 8
        qkv = qkv.bfloat16()
        o = torch.empty(B, T, D, dtype=torch.bfloat16, device=dev)
        lens = None
        if args.lens:
            lens = torch.tensor([int(x) for x in args.lens.split(",")], dtype=torch.int32, device=dev)

        def run():
            check(lib.rp_fmha(ptr(qkv), ptr(qkv) + D * 2, ptr(qkv) + 2 * D * 2, ptr(o), 3 * D, 3 * D, 3 * D, D,
                              T * 3 * D, T * 3 * D, T * 3 * D, T * D, B, H, T, T, ptr(lens), 0, 0, 0, 0,
                              cur_stream()), "fmha")
        ms = timeit(run, args.iters, flush)
        fl = 4.0 * B * H * T * T * 64
        out["fmha"] = {"ms": ms, "tflops": fl / ms / 1e9}
    if "gemm" in args.what:
        shapes = {"qkv(epi0)": (0, 1536, 512), "ff1(epi1)": (1, 2048, 512), "out(epi3)": (3, 512, 512),
                  "ff2(epi3)": (3, 512, 2048), "in(epi2)": (2, 512, 2944)}
        for name, (epi, N, K) in shapes.items():
            A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
            W = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
            bias = torch.randn(N, device=dev)
            f32 = epi >= 2
            Dm = torch.zeros(M, N, dtype=torch.float32 if f32 else torch.bfloat16, device=dev)

            def run():
                check(lib.rp_gemm_bf16(epi, ptr(A), K, ptr(W), K, ptr(Dm), N, ptr(bias),
                                       ptr(Dm) if epi == 3 else 0, N if epi == 3 else 0, M, N, K,
                                       cur_stream()), "gemm")
            ms = timeit(run, args.iters, flush)
            byt = M * K * 2 + N * K * 2 + M * N * (4 if f32 else 2) * (2 if epi == 3 else 1)
            out[f"gemm_{name}"] = {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9, "gbs": byt / ms / 1e6}
    if "ln" in args.what:
        x = torch.randn(M, 512, device=dev)
        g = torch.ones(512, device=dev)
        b = torch.zeros(512, device=dev)
        y = torch.empty(M, 512, dtype=torch.bfloat16, device=dev)

        def run():
            check(lib.rp_layernorm512(0, ptr(x), M, T, ptr(g), ptr(b), 0, 0, 0, 0, 0, 0, ptr(y), 0,
                                      cur_stream()), "ln")
        ms = timeit(run, args.iters, flush)
        out["ln_mode0"] = {"ms": ms, "gbs": M * 3072 / ms / 1e6}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
