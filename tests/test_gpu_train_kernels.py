"""Backward-pass kernels of the training step through the C ABI, each against torch's fp32
arithmetic for the same op (the reference's backward IS torch.autograd: main.py:326-333): dgrad / wgrad GEMMs with
MN-major operands and split-K, bias column sums, ReLU masks, the accumulating LayerNorm backward, the last cls-head
layer, the attention forward's log-sum-exp and the attention backward (dQ, dK, dV)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOG2E = 1.4426950408889634


def _lib():
    from repurpose_b200 import _lib
    return _lib, _lib.load()


def _scratch(lib):
    return torch.empty(int(lib.rp_train_scratch_bytes()), dtype=torch.uint8, device=DEV)


def _rel(got, ref):
    return ((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("M,N,K", [(300, 512, 512), (1000, 2048, 512), (57632, 512, 2048), (777, 256, 256), (2049, 512, 1536)])
@pytest.mark.parametrize("out_f32", [0, 1])
def test_gemm_dgrad(M, N, K, out_f32):
    """dX[M, N=in] = dY[M, K=out] W[K=out, N=in], the weight exactly as nn.Linear stores it."""
    L, lib = _lib()
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    dy = (torch.randn(M, K, device=DEV, generator=g) * 0.5).bfloat16()
    w = (torch.randn(K, N, device=DEV, generator=g) * 0.05).bfloat16()
    out = torch.full((M, N), float("nan"), dtype=torch.float32 if out_f32 else torch.bfloat16, device=DEV)
    L.check(lib.rp_gemm_bwd(1, out_f32, L.ptr(dy), K, L.ptr(w), N, L.ptr(out), N, M, N, K, 1, L.cur_stream()), "dgrad")
    ref = dy.float() @ w.float()
    assert _rel(out, ref) < (2e-5 if out_f32 else 6e-3), _rel(out, ref)


@pytest.mark.parametrize("tok,Nout,Kin,splits", [(1000, 512, 512, 1), (1801, 512, 2048, 4), (57632, 1536, 512, 6),
                                                 (57632, 512, 2944, 9), (333, 256, 256, 3), (4000, 2048, 512, 5)])
def test_gemm_wgrad_split_k(tok, Nout, Kin, splits):
    """dW[out, in] = dY[tok, out]^T X[tok, in]: both operands MN-major, reduction over any token count (TMA zero-fills the
    last k-block), column tail (2944 = 11.5 x 256), split-K partials reduced in a fixed order."""
    L, lib = _lib()
    g = torch.Generator(device=DEV).manual_seed(tok + Nout)
    dy = (torch.randn(tok, Nout, device=DEV, generator=g) * 0.5).bfloat16()
    x = (torch.randn(tok, Kin, device=DEV, generator=g) * 0.5).bfloat16()
    part = torch.full((splits * Nout, Kin), float("nan"), device=DEV)
    L.check(lib.rp_gemm_bwd(3, 1, L.ptr(dy), Nout, L.ptr(x), Kin, L.ptr(part), Kin, Nout, Kin, tok, splits, L.cur_stream()),
            "wgrad")
    dw = torch.empty(Nout, Kin, device=DEV)
    L.check(lib.rp_splitk_reduce(L.ptr(part), splits, Nout * Kin, L.ptr(dw), L.cur_stream()), "reduce")
    ref = dy.float().t() @ x.float()
    assert _rel(dw, ref) < 3e-5, _rel(dw, ref)
    part2 = torch.empty_like(part)
    L.check(lib.rp_gemm_bwd(3, 1, L.ptr(dy), Nout, L.ptr(x), Kin, L.ptr(part2), Kin, Nout, Kin, tok, splits, L.cur_stream()),
            "wgrad")
    assert torch.equal(part, part2)  # deterministic


@pytest.mark.parametrize("M,N", [(5, 256), (1000, 512), (57632, 1536), (4097, 2048)])
def test_colsum_and_relu_bwd(M, N):
    L, lib = _lib()
    g = torch.Generator(device=DEV).manual_seed(M)
    x = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    out = torch.empty(N, device=DEV)
    sc = _scratch(lib)
    L.check(lib.rp_colsum_bf16(L.ptr(x), M, N, L.ptr(out), L.ptr(sc), sc.numel(), L.cur_stream()), "colsum")
    ref = x.float().sum(0)
    assert torch.allclose(out, ref, atol=2e-3 * max(1.0, M ** 0.5), rtol=1e-4)
    act = torch.randn(M, N, device=DEV, generator=g).bfloat16()
    act[0, :8] = 0.0
    act[-1, -3:] = -0.0
    dy = x.clone()
    L.check(lib.rp_relu_bwd(L.ptr(dy), L.ptr(act), M * N, 0, L.cur_stream()), "relu_bwd")
    assert torch.equal(dy, torch.where(act > 0, x, torch.zeros_like(x)))
    dy2, cs = x.clone(), torch.empty(N, device=DEV)                              # the fused mask + bias gradient
    L.check(lib.rp_relu_bwd_colsum(L.ptr(dy2), L.ptr(act), M, N, L.ptr(cs), L.ptr(sc), sc.numel(), L.cur_stream()), "relu_colsum")
    assert torch.equal(dy2, dy)
    assert torch.allclose(cs, dy.float().sum(0), atol=2e-3 * max(1.0, M ** 0.5), rtol=1e-4)
    dy32, act32 = x.float().contiguous(), act.float().contiguous()
    exp32 = torch.where(act32 > 0, dy32, torch.zeros_like(dy32))
    L.check(lib.rp_relu_bwd(L.ptr(dy32), L.ptr(act32), M * N, 1, L.cur_stream()), "relu_bwd f32")
    assert torch.equal(dy32, exp32)


@pytest.mark.parametrize("M", [3, 1000, 57632])
def test_layernorm_bwd_accumulates_and_head_out_bwd(M):
    L, lib = _lib()
    g = torch.Generator(device=DEV).manual_seed(M)
    x = (torch.randn(M, 512, device=DEV, generator=g) * 2).requires_grad_(True)
    gamma = (torch.rand(512, device=DEV, generator=g) + 0.5).requires_grad_(True)
    beta = torch.zeros(512, device=DEV, requires_grad=True)
    dy = torch.randn(M, 512, device=DEV, generator=g)
    dh0 = torch.randn(M, 512, device=DEV, generator=g)
    torch.nn.functional.layer_norm(x, (512,), gamma, beta, 1e-5).backward(dy)
    dh = dh0.clone()
    dh16 = torch.empty(M, 512, dtype=torch.bfloat16, device=DEV)
    dgm, dbt = torch.empty(512, device=DEV), torch.empty(512, device=DEV)
    sc = _scratch(lib)
    csum = torch.empty(512, device=DEV)
    L.check(lib.rp_layernorm512_bwd_acc(L.ptr(x.detach()), L.ptr(dy), L.ptr(gamma.detach()), M, 1e-5, 1, L.ptr(dh), L.ptr(dh16),
                                        L.ptr(csum), L.ptr(dgm), L.ptr(dbt), L.ptr(sc), sc.numel(), L.cur_stream()), "ln_bwd_acc")
    assert torch.allclose(dh, dh0 + x.grad, atol=3e-5, rtol=1e-4)
    assert torch.equal(dh16, dh.bfloat16())
    tol = 2e-4 * max(1.0, M ** 0.5)
    assert torch.allclose(dgm, gamma.grad, atol=tol, rtol=1e-4) and torch.allclose(dbt, beta.grad, atol=tol, rtol=1e-4)
    assert torch.allclose(csum, dh.sum(0), atol=5 * tol, rtol=1e-4)          # bias gradient of the Linear in front
    dh2 = torch.full_like(dh0, float("nan"))                                   # accumulate = 0 overwrites
    L.check(lib.rp_layernorm512_bwd_acc(L.ptr(x.detach()), L.ptr(dy), L.ptr(gamma.detach()), M, 1e-5, 0, L.ptr(dh2), 0, 0,
                                        L.ptr(dgm), L.ptr(dbt), L.ptr(sc), sc.numel(), L.cur_stream()), "ln_bwd")
    assert torch.allclose(dh2, x.grad, atol=3e-5, rtol=1e-4)
    # last cls-head layer (Linear 256 -> 1) and the ReLU in front of it
    a2 = torch.relu(torch.randn(M, 256, device=DEV, generator=g)).bfloat16()
    w = torch.randn(256, device=DEV, generator=g)
    dlog = torch.randn(M, device=DEV, generator=g)
    da2 = torch.empty(M, 256, dtype=torch.bfloat16, device=DEV)
    dw, db = torch.empty(256, device=DEV), torch.empty(1, device=DEV)
    L.check(lib.rp_head_out_bwd(L.ptr(dlog), L.ptr(a2), L.ptr(w), M, L.ptr(da2), L.ptr(dw), L.ptr(db), L.ptr(sc), sc.numel(),
                                L.cur_stream()), "head_out_bwd")
    ref_da2 = torch.where(a2 > 0, dlog[:, None] * w[None, :], torch.zeros(M, 256, device=DEV))
    assert torch.allclose(da2.float(), ref_da2, atol=1e-2, rtol=1e-2)
    assert torch.allclose(dw, dlog @ a2.float(), atol=tol * 5, rtol=1e-3)
    assert torch.allclose(db, dlog.sum().reshape(1), atol=tol, rtol=1e-4)


def _attention_case(B, T, lens, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    H, D = 8, 512
    qkv = torch.randn(B, T, 3 * D, device=DEV, generator=g)
    qkv[..., :D] *= 1.5                     # some score spread
    qs = qkv.clone()
    qs[..., :D] *= LOG2E / 8                # the forward expects q pre-scaled by log2(e)/sqrt(d)
    d_o = torch.randn(B, T, D, device=DEV, generator=g).bfloat16()
    lens_t = torch.tensor(lens, dtype=torch.int32, device=DEV)
    return H, D, qkv.bfloat16(), qs.bfloat16(), d_o, lens_t


@pytest.fixture(params=["fused", "deterministic"])
def attn_bwd_mode(request):
    """both attention-backward paths: the fused kernel (dQ summed by TMA reduce adds in L2, the default) and the two
    deterministic kernels (rp_set_attn_bwd_deterministic)"""
    L, lib = _lib()
    L.check(lib.rp_set_attn_bwd_deterministic(1 if request.param == "deterministic" else 0), "set mode")
    yield request.param
    L.check(lib.rp_set_attn_bwd_deterministic(-1), "set mode")


@pytest.mark.parametrize("B,T,lens", [(2, 40, [40, 17]), (1, 128, [128]), (2, 200, [200, 77]), (2, 333, [333, 129]),
                                      (1, 700, [641]), (3, 1801, [1801, 1211, 64])])
def test_fmha_backward_matches_autograd(B, T, lens, attn_bwd_mode):
    L, lib = _lib()
    H, D, qkv, qs, d_o, lens_t = _attention_case(B, T, lens, 17 * T + B)
    o = torch.empty(B, T, D, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(B, H, T, device=DEV)
    L.check(lib.rp_fmha_train(L.ptr(qs), L.ptr(qs) + 2 * D, L.ptr(qs) + 4 * D, L.ptr(o), 3 * D, D, B, H, T, L.ptr(lens_t),
                              L.ptr(lse), L.cur_stream()), "fmha_train")
    # fp32 reference on the bf16-rounded inputs (q = q' 8 / log2e is what the scaled operand stands for)
    q = (qs[..., :D].float() * (8 / LOG2E)).view(B, T, H, 64).transpose(1, 2).requires_grad_(True)
    k = qs[..., D:2 * D].float().view(B, T, H, 64).transpose(1, 2).requires_grad_(True)
    v = qs[..., 2 * D:].float().view(B, T, H, 64).transpose(1, 2).requires_grad_(True)
    s = (q @ k.transpose(-1, -2)) / 8
    keymask = torch.arange(T, device=DEV)[None, :] >= lens_t[:, None]
    s = s.masked_fill(keymask[:, None, None, :], float("-inf"))
    p_ref = torch.softmax(s, -1)
    o_ref = (p_ref @ v).transpose(1, 2).reshape(B, T, D)
    assert _rel(o, o_ref) < 2e-2
    lse_ref = torch.logsumexp(s, -1) * LOG2E
    assert torch.allclose(lse, lse_ref, atol=2e-2, rtol=1e-3), (lse - lse_ref).abs().max().item()
    o_ref.backward(d_o.float())
    dqkv = torch.full((B, T, 3 * D), float("nan"), dtype=torch.bfloat16, device=DEV)
    dsum = torch.empty(B, H, T, device=DEV)
    L.check(lib.rp_fmha_bwd(L.ptr(qs), L.ptr(qs) + 2 * D, L.ptr(qs) + 4 * D, L.ptr(o), L.ptr(d_o), L.ptr(lse), L.ptr(dsum),
                            L.ptr(dqkv), L.ptr(dqkv) + 2 * D, L.ptr(dqkv) + 4 * D, 3 * D, D, 3 * D, B, H, T, L.ptr(lens_t),
                            L.cur_stream()), "fmha_bwd")
    torch.cuda.synchronize()
    assert not torch.isnan(dqkv.float()).any()
    for name, got, ref in (("dq", dqkv[..., :D], q.grad), ("dk", dqkv[..., D:2 * D], k.grad), ("dv", dqkv[..., 2 * D:], v.grad)):
        ref = ref.transpose(1, 2).reshape(B, T, D)
        err = _rel(got, ref)
        assert err < 3e-2, (name, err)
    for b, n in enumerate(lens):   # padded keys receive exactly zero gradient
        assert (dqkv[b, n:, D:].float() == 0).all()
    # a second launch: bit-identical with the deterministic kernels; with the fused kernel dK / dV are, and dQ agrees to
    # the rounding of its fp32 sums over the key tiles (their order in L2 is not fixed)
    again = torch.full_like(dqkv, float("nan"))
    L.check(lib.rp_fmha_bwd(L.ptr(qs), L.ptr(qs) + 2 * D, L.ptr(qs) + 4 * D, L.ptr(o), L.ptr(d_o), L.ptr(lse), L.ptr(dsum),
                            L.ptr(again), L.ptr(again) + 2 * D, L.ptr(again) + 4 * D, 3 * D, D, 3 * D, B, H, T, L.ptr(lens_t),
                            L.cur_stream()), "fmha_bwd")
    torch.cuda.synchronize()
    assert torch.equal(again[..., D:], dqkv[..., D:])
    if attn_bwd_mode == "deterministic":
        assert torch.equal(again, dqkv)
    else:
        assert _rel(again[..., :D], dqkv[..., :D]) < 1e-2


def test_fmha_backward_of_a_video_without_keys(attn_bwd_mode):
    """kv_len = 0 (a slot of a fixed-shape batch that holds no video): the forward writes lse = +inf and O = 0, the backward
    exact zeros for that batch element — in both modes, and without touching its neighbour"""
    L, lib = _lib()
    B, T, lens = 2, 200, [200, 0]
    H, D, qkv, qs, d_o, lens_t = _attention_case(B, T, lens, 23)
    o = torch.full((B, T, D), float("nan"), dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(B, H, T, device=DEV)
    L.check(lib.rp_fmha_train(L.ptr(qs), L.ptr(qs) + 2 * D, L.ptr(qs) + 4 * D, L.ptr(o), 3 * D, D, B, H, T, L.ptr(lens_t),
                              L.ptr(lse), L.cur_stream()), "fmha_train")
    dqkv = torch.full((B, T, 3 * D), float("nan"), dtype=torch.bfloat16, device=DEV)
    dsum = torch.empty(B, H, T, device=DEV)
    L.check(lib.rp_fmha_bwd(L.ptr(qs), L.ptr(qs) + 2 * D, L.ptr(qs) + 4 * D, L.ptr(o), L.ptr(d_o), L.ptr(lse), L.ptr(dsum),
                            L.ptr(dqkv), L.ptr(dqkv) + 2 * D, L.ptr(dqkv) + 4 * D, 3 * D, D, 3 * D, B, H, T, L.ptr(lens_t),
                            L.cur_stream()), "fmha_bwd")
    torch.cuda.synchronize()
    assert (o[1].float() == 0).all() and (dqkv[1].float() == 0).all()
    assert not torch.isnan(dqkv[0].float()).any() and float(dqkv[0].float().abs().max()) > 0


def test_fmha_backward_fused_agrees_with_the_deterministic_kernels():
    """same inputs through both paths: dK and dV bit-identical (the same bf16 dSt^T / P^T tiles feed the same MMAs), dQ equal
    up to the bf16 rounding of differently ordered fp32 sums"""
    L, lib = _lib()
    B, T, lens = 2, 1801, [1801, 700]
    H, D, qkv, qs, d_o, lens_t = _attention_case(B, T, lens, 5)
    o = torch.empty(B, T, D, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(B, H, T, device=DEV)
    L.check(lib.rp_fmha_train(L.ptr(qs), L.ptr(qs) + 2 * D, L.ptr(qs) + 4 * D, L.ptr(o), 3 * D, D, B, H, T, L.ptr(lens_t),
                              L.ptr(lse), L.cur_stream()), "fmha_train")
    out = {}
    try:
        for mode in (0, 1):
            L.check(lib.rp_set_attn_bwd_deterministic(mode), "set mode")
            dqkv = torch.full((B, T, 3 * D), float("nan"), dtype=torch.bfloat16, device=DEV)
            dsum = torch.empty(B, H, T, device=DEV)
            L.check(lib.rp_fmha_bwd(L.ptr(qs), L.ptr(qs) + 2 * D, L.ptr(qs) + 4 * D, L.ptr(o), L.ptr(d_o), L.ptr(lse),
                                    L.ptr(dsum), L.ptr(dqkv), L.ptr(dqkv) + 2 * D, L.ptr(dqkv) + 4 * D, 3 * D, D, 3 * D, B, H, T,
                                    L.ptr(lens_t), L.cur_stream()), "fmha_bwd")
            torch.cuda.synchronize()
            out[mode] = dqkv
    finally:
        L.check(lib.rp_set_attn_bwd_deterministic(-1), "set mode")
    assert not torch.isnan(out[0].float()).any()
    assert torch.equal(out[0][..., D:], out[1][..., D:])
    err = _rel(out[0][..., :D], out[1][..., :D])
    print("dQ fused vs deterministic: relative error", err)
    assert err < 1e-2


def test_gemm_wgrad_refuses_an_empty_split():
    """10 k-blocks in 6 splits of 2 would leave the sixth split without work (its accumulator barrier would never
    complete): the launcher refuses instead of hanging."""
    L, lib = _lib()
    dy = torch.zeros(640, 512, dtype=torch.bfloat16, device=DEV)
    x = torch.zeros(640, 512, dtype=torch.bfloat16, device=DEV)
    part = torch.zeros(6 * 512, 512, device=DEV)
    rc = lib.rp_gemm_bwd(3, 1, L.ptr(dy), 512, L.ptr(x), 512, L.ptr(part), 512, 512, 512, 640, 6, L.cur_stream())
    assert rc != 0 and b"empty" in lib.rp_last_error()
    L.check(lib.rp_gemm_bwd(3, 1, L.ptr(dy), 512, L.ptr(x), 512, L.ptr(part), 512, 512, 512, 640, 5, L.cur_stream()), "wgrad")
