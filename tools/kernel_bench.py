#!/usr/bin/env python
"""Micro-benchmark of the individual kernels at the bench.py shapes (B=32, T=1801), timed with CUDA
events.  Used while optimising and as the target command for `ncu -k regex:<kernel>`.

    python tools/kernel_bench.py [fmha] [gemm] [gemmln] [rowwise] [ln] [--iters N] [--B 32] [--T 1801]
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
    if "fmhabwd" in args.what:
        qkv = torch.randn(B, T, 3 * D, device=dev)
        qkv[..., :D] *= LOG2E / 8
        qkv = qkv.bfloat16()
        o = torch.empty(B, T, D, dtype=torch.bfloat16, device=dev)
        d_o = torch.randn(B, T, D, device=dev).bfloat16()
        dqkv = torch.empty(B, T, 3 * D, dtype=torch.bfloat16, device=dev)
        lse = torch.empty(B, H, T, device=dev)
        dsum = torch.empty(B, H, T, device=dev)
        lens = torch.full((B,), T, dtype=torch.int32, device=dev)

        def fwd():
            check(lib.rp_fmha_train(ptr(qkv), ptr(qkv) + D * 2, ptr(qkv) + 2 * D * 2, ptr(o), 3 * D, D, B, H, T, ptr(lens),
                                    ptr(lse), cur_stream()), "fmha_train")

        def bwd():
            check(lib.rp_fmha_bwd(ptr(qkv), ptr(qkv) + D * 2, ptr(qkv) + 2 * D * 2, ptr(o), ptr(d_o), ptr(lse), ptr(dsum),
                                  ptr(dqkv), ptr(dqkv) + D * 2, ptr(dqkv) + 2 * D * 2, 3 * D, D, 3 * D, B, H, T, ptr(lens),
                                  cur_stream()), "fmha_bwd")
        ms = timeit(fwd, args.iters, flush)
        out["fmha_train_fwd"] = {"ms": ms, "tflops": 4.0 * B * H * T * T * 64 / ms / 1e9}
        ms = timeit(bwd, args.iters, flush)
        # algorithmic: 5 T x T x 64 contractions (dV, dP, dQ, dK + the S recompute); the two-kernel split executes 7
        out["fmha_bwd"] = {"ms": ms, "tflops_algorithmic": 10.0 * B * H * T * T * 64 / ms / 1e9,
                           "tflops_executed": 14.0 * B * H * T * T * 64 / ms / 1e9}
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
    if "gemmln" in args.what:
        # the LayerNorm-fused residual GEMMs of the forward (out-proj K=512, FF2 K=2048): h += A W^T + b; u = LN(h)
        for name, K in (("out+ln", 512), ("ff2+ln", 2048)):
            A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
            W = (torch.randn(512, K, device=dev) * 0.05).bfloat16()
            bias = torch.randn(512, device=dev)
            hres = torch.zeros(M, 512, device=dev)
            u = torch.empty(M, 512, dtype=torch.bfloat16, device=dev)
            g = torch.ones(512, device=dev)
            b = torch.zeros(512, device=dev)

            def run():
                check(lib.rp_gemm_resid_ln(ptr(A), K, ptr(W), K, ptr(hres), 512, ptr(bias), ptr(g), ptr(b), 1e-5,
                                           ptr(u), 512, M, K, cur_stream()), "gemm_resid_ln")
            ms = timeit(run, args.iters, flush)
            byt = M * (K * 2 + 512 * 4 * 2 + 512 * 2) + 512 * K * 2
            out[f"gemm_{name}"] = {"ms": ms, "tflops": 2.0 * M * 512 * K / ms / 1e9, "gbs": byt / ms / 1e6}
    if "rowwise" in args.what:
        # memory-bound kernels of the step, against their algorithmic bytes
        a_c = torch.randn(M, 256, device=dev).bfloat16()
        a_r = torch.randn(M, 256, device=dev).bfloat16()
        wc, bc = torch.randn(256, device=dev), torch.randn(1, device=dev)
        wr, br = torch.randn(2, 256, device=dev), torch.randn(2, device=dev)
        lg = torch.empty(M, device=dev)
        off = torch.empty(M, 2, device=dev)

        def run():
            check(lib.rp_head_out(ptr(a_c), ptr(a_r), ptr(wc), ptr(bc), ptr(wr), ptr(br), ptr(lg), ptr(off), M,
                                  cur_stream()), "head_out")
        ms = timeit(run, args.iters, flush)
        out["head_out"] = {"ms": ms, "gbs": M * (2 * 512 + 12) / ms / 1e6}
        x = torch.randn(M, 512, device=dev)
        g = torch.ones(512, device=dev)
        b = torch.zeros(512, device=dev)
        pe = torch.randn(T, 512, device=dev)
        hf = torch.empty(M, 512, device=dev)
        y = torch.empty(M, 512, dtype=torch.bfloat16, device=dev)
        y2 = torch.empty(M, 512, dtype=torch.bfloat16, device=dev)

        def run():
            check(lib.rp_layernorm512(1, ptr(x), M, T, ptr(g), ptr(b), ptr(g), ptr(b), 0, 0, ptr(pe), ptr(hf), ptr(y), 0,
                                      cur_stream()), "ln1")
        ms = timeit(run, args.iters, flush)
        out["ln_mode1"] = {"ms": ms, "gbs": M * 5120 / ms / 1e6}

        def run():
            check(lib.rp_layernorm512(2, ptr(x), M, T, ptr(g), ptr(b), ptr(g), ptr(b), ptr(g), ptr(b), 0, ptr(hf), ptr(y),
                                      ptr(y2), cur_stream()), "ln2")
        ms = timeit(run, args.iters, flush)
        out["ln_mode2"] = {"ms": ms, "gbs": M * 6144 / ms / 1e6}
        vis, aud, txt = (torch.randn(M, c, device=dev) for c in (512, 2048, 384))
        xc = torch.empty(M, 2944, dtype=torch.bfloat16, device=dev)

        def run():
            check(lib.rp_concat_cast(ptr(vis), ptr(aud), ptr(txt), 512, 2048, 384, ptr(xc), M, cur_stream()), "concat")
        ms = timeit(run, args.iters, flush)
        out["concat_cast"] = {"ms": ms, "gbs": M * 2944 * 6 / ms / 1e6}
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
