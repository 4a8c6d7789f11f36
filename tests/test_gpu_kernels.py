"""GPU parity tests of the individual sm_100a kernels, driven through the C ABI (ctypes).
Each kernel is compared with a plain PyTorch fp32 evaluation of the same op on the same
(bf16-rounded where the kernel consumes bf16) inputs."""
import math

import numpy as np
import pytest
import torch

from repurpose_b200 import _lib
from repurpose_b200._lib import check, cur_stream, ptr

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOG2E = 1.4426950408889634


def _lib_():
    return _lib.load()


def _describe(got, ref, name):
    d = (got.float() - ref.float()).abs()
    bad = d > (1e-2 + 2e-2 * ref.float().abs())
    msg = (f"{name}: max|diff|={d.max().item():.4g} at {np.unravel_index(int(d.argmax()), d.shape)}, "
           f"max|ref|={ref.float().abs().max().item():.4g}, mismatching={bad.float().mean().item():.3%}")
    if bad.any() and d.dim() == 2:
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        msg += (f"; bad rows {rows[:8].tolist()}..({rows.numel()}) bad cols {cols[:8].tolist()}..({cols.numel()})"
                f"; got[0,:4]={got[0, :4].tolist()} ref[0,:4]={ref[0, :4].tolist()}")
    return msg


# ------------------------------------------------------------------------------------------ GEMM
def _gemm(epi, A, W, bias, resid=None, out_dtype=torch.bfloat16):
    M, K = A.shape
    N = W.shape[0]
    D = resid if (epi == 3 and resid is not None) else torch.empty(M, N, dtype=out_dtype, device=DEV)
    check(_lib_().rp_gemm_bf16(epi, ptr(A), A.stride(0), ptr(W), W.stride(0), ptr(D), D.stride(0),
                               ptr(bias), ptr(resid), 0 if resid is None else resid.stride(0), M, N, K,
                               cur_stream()), "rp_gemm_bf16")
    torch.cuda.synchronize()
    return D


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 512), (300, 512, 512), (1000, 1536, 512),
                                   (4000, 2048, 512), (777, 512, 2048), (2048, 512, 2944), (64, 256, 256),
                                   (57632, 512, 512)])
def test_gemm_bias_bf16(M, N, K):
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    A = (torch.randn(M, K, device=DEV, generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g)
    ref = A.float() @ W.float().T + bias
    got = _gemm(0, A, W, bias)
    assert torch.allclose(got.float(), ref, atol=2e-2, rtol=2e-2), _describe(got, ref, f"gemm {M}x{N}x{K}")


def test_gemm_epilogues():
    M, N, K = 1000, 512, 512
    g = torch.Generator(device=DEV).manual_seed(7)
    A = (torch.randn(M, K, device=DEV, generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g)
    acc = A.float() @ W.float().T + bias
    got = _gemm(1, A, W, bias)
    assert torch.allclose(got.float(), acc.relu(), atol=2e-2, rtol=2e-2), _describe(got, acc.relu(), "relu")
    got = _gemm(2, A, W, bias, out_dtype=torch.float32)
    assert torch.allclose(got, acc, atol=2e-3, rtol=2e-3), _describe(got, acc, "f32 out")
    resid = torch.randn(M, N, device=DEV, generator=g)
    want = acc + resid
    got = _gemm(3, A, W, bias, resid=resid.clone(), out_dtype=torch.float32)  # in place
    assert torch.allclose(got, want, atol=2e-3, rtol=2e-3), _describe(got, want, "residual")


def test_gemm_strided_operands():
    # A is a column slice of a wider buffer (lda > K), output written into a wider buffer (ldd > N)
    M, N, K = 500, 256, 256
    g = torch.Generator(device=DEV).manual_seed(3)
    big = (torch.randn(M, 1024, device=DEV, generator=g) * 0.5).bfloat16()
    A = big[:, 256:512]
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.05).bfloat16()
    bias = torch.zeros(N, device=DEV)
    out = torch.zeros(M, 768, dtype=torch.bfloat16, device=DEV)
    D = out[:, 256:512]
    check(_lib_().rp_gemm_bf16(0, ptr(A), 1024, ptr(W), K, ptr(D), 768, ptr(bias), 0, 0, M, N, K,
                               cur_stream()), "gemm")
    torch.cuda.synchronize()
    ref = A.float() @ W.float().T
    assert torch.allclose(D.float(), ref, atol=2e-2, rtol=2e-2), _describe(D, ref, "strided")
    assert out[:, :256].abs().max() == 0 and out[:, 512:].abs().max() == 0, "wrote outside the tile"


def test_gemm_rejects_bad_shapes():
    A = torch.zeros(8, 64, dtype=torch.bfloat16, device=DEV)
    W = torch.zeros(100, 64, dtype=torch.bfloat16, device=DEV)
    D = torch.zeros(8, 100, dtype=torch.bfloat16, device=DEV)
    rc = _lib_().rp_gemm_bf16(0, ptr(A), 64, ptr(W), 64, ptr(D), 100, 0, 0, 0, 8, 100, 64, cur_stream())
    assert rc == 1 and b"multiple" in _lib_().rp_last_error()


# ------------------------------------------------------------------------------- row-wise kernels
def test_concat_cast():
    M = 777
    v, a, t = (torch.randn(M, c, device=DEV) for c in (512, 2048, 384))
    out = torch.empty(M, 2944, dtype=torch.bfloat16, device=DEV)
    check(_lib_().rp_concat_cast(ptr(v), ptr(a), ptr(t), 512, 2048, 384, ptr(out), M, cur_stream()), "cc")
    torch.cuda.synchronize()
    assert torch.equal(out, torch.cat([v, a, t], -1).bfloat16())


def _ln(x, g, b):
    return torch.nn.functional.layer_norm(x, (512,), g, b, 1e-5)


def test_layernorm_modes():
    M, T = 1003, 59
    gen = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(M, 512, device=DEV, generator=gen) * 3 + 1
    ps = [torch.randn(512, device=DEV, generator=gen) for _ in range(6)]
    pe = torch.randn(T, 512, device=DEV, generator=gen)
    lib = _lib_()
    y = torch.empty(M, 512, dtype=torch.bfloat16, device=DEV)
    y2 = torch.empty_like(y)
    f = torch.empty(M, 512, device=DEV)
    check(lib.rp_layernorm512(0, ptr(x), M, T, ptr(ps[0]), ptr(ps[1]), 0, 0, 0, 0, 0, 0, ptr(y), 0,
                              cur_stream()), "ln0")
    torch.cuda.synchronize()
    ref = _ln(x, ps[0], ps[1])
    assert torch.allclose(y.float(), ref, atol=3e-2, rtol=1e-2), _describe(y, ref, "ln mode 0")
    check(lib.rp_layernorm512(3, ptr(x), M, T, ptr(ps[0]), ptr(ps[1]), 0, 0, 0, 0, 0, ptr(f), 0, 0,
                              cur_stream()), "ln3")
    torch.cuda.synchronize()
    assert torch.allclose(f, ref, atol=1e-5, rtol=1e-5), _describe(f, ref, "ln mode 3")
    # mode 1 in place: h = LN(x)+pe[row % T]; y = LN1(h)
    xin = x.clone()
    check(lib.rp_layernorm512(1, ptr(xin), M, T, ptr(ps[0]), ptr(ps[1]), ptr(ps[2]), ptr(ps[3]), 0, 0,
                              ptr(pe), ptr(xin), ptr(y), 0, cur_stream()), "ln1")
    torch.cuda.synchronize()
    h = ref + pe[torch.arange(M, device=DEV) % T]
    assert torch.allclose(xin, h, atol=1e-5, rtol=1e-5), _describe(xin, h, "ln mode 1 h")
    r1 = _ln(h, ps[2], ps[3])
    assert torch.allclose(y.float(), r1, atol=3e-2, rtol=1e-2), _describe(y, r1, "ln mode 1 y")
    # mode 2
    check(lib.rp_layernorm512(2, ptr(x), M, T, ptr(ps[0]), ptr(ps[1]), ptr(ps[2]), ptr(ps[3]), ptr(ps[4]),
                              ptr(ps[5]), 0, ptr(f), ptr(y), ptr(y2), cur_stream()), "ln2")
    torch.cuda.synchronize()
    fr = ref.relu()
    assert torch.allclose(f, fr, atol=1e-5, rtol=1e-5)
    assert torch.allclose(y.float(), _ln(fr, ps[2], ps[3]), atol=3e-2, rtol=1e-2)
    assert torch.allclose(y2.float(), _ln(fr, ps[4], ps[5]), atol=3e-2, rtol=1e-2)


def test_head_out():
    M = 999
    gen = torch.Generator(device=DEV).manual_seed(2)
    ac = torch.randn(M, 256, device=DEV, generator=gen).bfloat16()
    ar = torch.randn(M, 256, device=DEV, generator=gen).bfloat16()
    wc, bc = torch.randn(1, 256, device=DEV, generator=gen), torch.randn(1, device=DEV, generator=gen)
    wr, br = torch.randn(2, 256, device=DEV, generator=gen), torch.randn(2, device=DEV, generator=gen)
    logits = torch.empty(M, device=DEV)
    offs = torch.empty(M, 2, device=DEV)
    check(_lib_().rp_head_out(ptr(ac), ptr(ar), ptr(wc), ptr(bc), ptr(wr), ptr(br), ptr(logits), ptr(offs),
                              M, cur_stream()), "head_out")
    torch.cuda.synchronize()
    assert torch.allclose(logits, (ac.float() @ wc.T + bc).squeeze(1), atol=1e-3, rtol=1e-4)
    assert torch.allclose(offs, (ar.float() @ wr.T + br).relu(), atol=1e-3, rtol=1e-4)


@pytest.mark.parametrize("M", [7, 128, 129, 1000, 57632])
@pytest.mark.parametrize("nj,final_relu", [(1, False), (2, True)])
def test_gemm_head_dot_fuses_the_last_two_head_layers(M, nj, final_relu):
    """Linear(256,256) + ReLU + Linear(256, nj) (+ ReLU): the tail of cls_head / reg_head (models/MMCTransformer.py:71-93)"""
    gen = torch.Generator(device=DEV).manual_seed(M + nj)
    a1 = torch.relu(torch.randn(M, 256, device=DEV, generator=gen)).bfloat16()
    w4 = (torch.randn(256, 256, device=DEV, generator=gen) * 0.08).bfloat16()
    b4 = torch.randn(256, device=DEV, generator=gen) * 0.1
    w7 = torch.randn(nj, 256, device=DEV, generator=gen) * 0.1
    b7 = torch.randn(nj, device=DEV, generator=gen)
    out = torch.full((M, nj), float("nan"), device=DEV)
    check(_lib_().rp_gemm_head_dot(ptr(a1), 256, ptr(w4), 256, ptr(b4), ptr(w7), ptr(b7), nj, int(final_relu), ptr(out), M, 256,
                                   cur_stream()), "gemm_head_dot")
    ref = torch.relu(a1.float() @ w4.float().T + b4) @ w7.T + b7
    if final_relu:
        ref = ref.relu()
    assert torch.allclose(out, ref, atol=2e-3, rtol=1e-3), (out - ref).abs().max().item()


# ------------------------------------------------------------------------------------------ FMHA
def _fmha(q, k, v, kv_lens=None, mask=None):
    """q [B,Tq,H*64] bf16 (already scaled), k/v [B,Tk,H*64] bf16 (possibly column views)."""
    B, Tq, D = q.shape
    Tk = k.shape[1]
    H = D // 64
    o = torch.zeros(B, Tq, D, dtype=torch.bfloat16, device=DEV)
    mode, mb, mq, mptr = 0, 0, 0, 0
    if mask is not None:
        mask = mask.to(torch.uint8).contiguous()
        mode, mb, mptr = 1, mask.stride(0), ptr(mask)
        mq = 0 if mask.shape[1] == 1 else mask.stride(1)
    check(_lib_().rp_fmha(ptr(q), ptr(k), ptr(v), ptr(o), q.stride(1), k.stride(1), v.stride(1),
                          o.stride(1), q.stride(0), k.stride(0), v.stride(0), o.stride(0), B, H, Tq, Tk,
                          ptr(kv_lens), mode, mptr, mb, mq, cur_stream()), "rp_fmha")
    torch.cuda.synchronize()
    return o


def _attn_ref(q, k, v, kv_lens=None, mask=None):
    """fp32 reference; q is pre-scaled by log2e/8, so undo log2e and use exp."""
    B, Tq, D = q.shape
    H = D // 64
    qf = q.float().view(B, Tq, H, 64).transpose(1, 2) / LOG2E
    kf = k.float().view(B, -1, H, 64).transpose(1, 2)
    vf = v.float().view(B, -1, H, 64).transpose(1, 2)
    s = qf @ kf.transpose(-1, -2)
    if kv_lens is not None:
        pad = torch.arange(k.shape[1], device=DEV)[None, :] >= kv_lens[:, None]
        s = s.masked_fill(pad[:, None, None, :], float("-inf"))
    if mask is not None:
        s = s.masked_fill(mask[:, None] == 0, -1e9)
    return (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B, Tq, D)


@pytest.mark.parametrize("B,H,T,lens", [(1, 1, 128, None), (1, 1, 256, None), (2, 2, 384, None),
                                        (2, 8, 200, [200, 77]), (3, 8, 700, [700, 433, 129]),
                                        (2, 8, 1801, [1801, 1200]), (1, 8, 130, [1]),
                                        # query blocks of 256 rows: partial second tile (1, 2, 3 warps), a
                                        # one-row block; key tails of exactly 16 / 64 / 65 / 128 keys
                                        (1, 2, 300, None), (2, 2, 416, [416, 300]), (1, 1, 257, None),
                                        (1, 2, 512, [144]), (2, 2, 640, [64, 65]), (2, 1, 352, [128, 16]),
                                        (1, 1, 16, None)])
def test_fmha_key_padding(B, H, T, lens):
    gen = torch.Generator(device=DEV).manual_seed(T + B)
    qkv = torch.randn(B, T, 3 * H * 64, device=DEV, generator=gen)
    qkv[..., :H * 64] *= LOG2E / 8.0
    qkv = qkv.bfloat16()
    q, k, v = qkv[..., :H * 64], qkv[..., H * 64:2 * H * 64], qkv[..., 2 * H * 64:]
    kv = None if lens is None else torch.tensor(lens, dtype=torch.int32, device=DEV)
    got = _fmha(q, k, v, kv)
    ref = _attn_ref(q, k, v, kv)
    g2, r2 = got.view(B * T, -1), ref.view(B * T, -1)
    assert torch.isfinite(got.float()).all(), "non-finite attention output"
    assert torch.allclose(g2.float(), r2, atol=2e-2, rtol=2e-2), _describe(g2, r2, f"fmha B{B} H{H} T{T}")


def test_fmha_peaked_scores_rescale_path():
    # large-magnitude logits whose running max keeps increasing along the key axis exercise the
    # O-rescale (correction) branch
    B, H, T = 1, 2, 1024
    gen = torch.Generator(device=DEV).manual_seed(5)
    q = torch.randn(B, T, H * 64, device=DEV, generator=gen)
    k = torch.randn(B, T, H * 64, device=DEV, generator=gen)
    k = k * torch.linspace(0.2, 6.0, T, device=DEV)[None, :, None]
    v = torch.randn(B, T, H * 64, device=DEV, generator=gen)
    q = (q * LOG2E / 8.0).bfloat16()
    k, v = k.bfloat16(), v.bfloat16()
    got = _fmha(q, k, v)
    ref = _attn_ref(q, k, v)
    g2, r2 = got.view(B * T, -1), ref.view(B * T, -1)
    assert torch.allclose(g2.float(), r2, atol=3e-2, rtol=3e-2), _describe(g2, r2, "fmha peaked")


@pytest.mark.parametrize("Tq,Tk,kind", [(200, 200, "band"), (150, 333, "padding"), (300, 129, "random"),
                                        (64, 64, "fully_masked_row")])
def test_fmha_explicit_mask(Tq, Tk, kind):
    B, H = 2, 8
    gen = torch.Generator(device=DEV).manual_seed(Tq * 7 + Tk)
    q = (torch.randn(B, Tq, H * 64, device=DEV, generator=gen) * LOG2E / 8.0).bfloat16()
    k = torch.randn(B, Tk, H * 64, device=DEV, generator=gen).bfloat16()
    v = torch.randn(B, Tk, H * 64, device=DEV, generator=gen).bfloat16()
    if kind == "band":
        i = torch.arange(Tq, device=DEV)[:, None]
        j = torch.arange(Tk, device=DEV)[None, :]
        mask = ((i - j).abs() <= 16).expand(B, Tq, Tk)
    elif kind == "padding":
        lens = torch.tensor([Tk, Tk // 3], device=DEV)
        mask = (torch.arange(Tk, device=DEV)[None, :] < lens[:, None])[:, None, :]
    elif kind == "random":
        mask = torch.rand(B, Tq, Tk, device=DEV, generator=gen) > 0.5
    else:
        mask = torch.ones(B, Tq, Tk, dtype=torch.bool, device=DEV)
        mask[:, 5] = False  # fully masked query row -> uniform attention (reference: -1e9, not -inf)
    got = _fmha(q, k, v, mask=mask)
    ref = _attn_ref(q, k, v, mask=mask.to(torch.uint8))
    g2, r2 = got.view(B * Tq, -1), ref.view(B * Tq, -1)
    assert torch.allclose(g2.float(), r2, atol=2e-2, rtol=2e-2), _describe(g2, r2, f"fmha mask {kind}")


@pytest.mark.parametrize("M,K", [(129, 512), (256, 512), (257, 2048), (300, 512), (1000, 2048), (4099, 2048),
                                 (8448, 512), (8449, 512), (16897, 2048), (57632, 512)])
def test_gemm_residual_layernorm_fused(M, K):
    """h += A W^T + b and u = LayerNorm(h) in one kernel (clusters of four CTAs exchange row statistics)."""
    lib = _lib.load()
    torch.manual_seed(M + K)
    A = (torch.randn(M, K, device=DEV) * 0.5).bfloat16()
    W = (torch.randn(512, K, device=DEV) / K ** 0.5).bfloat16()
    bias = torch.randn(512, device=DEV)
    h0 = torch.randn(M, 512, device=DEV) * 2 + torch.randn(M, 1, device=DEV)       # rows with a mean offset
    gamma, beta = torch.rand(512, device=DEV) + 0.5, torch.randn(512, device=DEV)
    h = h0.clone()
    u = torch.empty(M, 512, device=DEV, dtype=torch.bfloat16)
    check(lib.rp_gemm_resid_ln(ptr(A), K, ptr(W), K, ptr(h), 512, ptr(bias), ptr(gamma), ptr(beta), 1e-5, ptr(u), 512,
                               M, K, cur_stream()), "rp_gemm_resid_ln")
    ref_h = h0 + A.float() @ W.float().t() + bias
    ref_u = torch.nn.functional.layer_norm(ref_h, (512,), gamma, beta, 1e-5)
    assert torch.allclose(h, ref_h, atol=2e-3, rtol=1e-4), (h - ref_h).abs().max().item()
    err = (u.float() - ref_u).abs().max().item()
    assert err < 3e-2, err                      # bf16 output of O(1..4) values
    # the unfused pair of kernels computes the same thing: h bit for bit, u within bf16 rounding of the statistics
    h2 = h0.clone()
    check(lib.rp_gemm_bf16(3, ptr(A), K, ptr(W), K, ptr(h2), 512, ptr(bias), ptr(h2), 512, M, 512, K, cur_stream()), "gemm")
    assert torch.equal(h, h2)


def test_gemm_fused_layernorm_survives_a_large_row_offset():
    """residual rows with a large common offset (mean^2 >> var, as in trained checkpoints with
    massive activations) must not lose the variance to cancellation: the fused epilogue accumulates shifted
    statistics and merges the two 256-column halves with the pairwise (Chan) update."""
    lib = _lib.load()
    torch.manual_seed(7)
    M, K = 1000, 512
    A = (torch.randn(M, K, device=DEV) * 0.5).bfloat16()
    W = (torch.randn(512, K, device=DEV) / K ** 0.5).bfloat16()
    bias = torch.randn(512, device=DEV)
    gamma, beta = torch.rand(512, device=DEV) + 0.5, torch.randn(512, device=DEV)
    for offset in (100.0, -3000.0):
        h0 = torch.randn(M, 512, device=DEV) + offset
        h0[:, 256:] += 0.5                                   # the two halves of a row get different means
        h = h0.clone()
        u = torch.empty(M, 512, device=DEV, dtype=torch.bfloat16)
        check(lib.rp_gemm_resid_ln(ptr(A), K, ptr(W), K, ptr(h), 512, ptr(bias), ptr(gamma), ptr(beta), 1e-5, ptr(u),
                                   512, M, K, cur_stream()), "rp_gemm_resid_ln")
        ref_h = (h0.double() + A.double() @ W.double().t() + bias.double())
        ref_u = torch.nn.functional.layer_norm(ref_h, (512,), gamma.double(), beta.double(), 1e-5).float()
        err = (u.float() - ref_u).abs().max().item()
        assert err < 4e-2, (offset, err)                     # bf16 output of O(1..4) values (fp32 h carries ~1e-4 of noise at |h| ~ 3000)
