"""Train-mode dropout of the training step (the reference trains under model.train(), main.py:285, with
nn.Dropout(0.1) in nn.TransformerEncoderLayer, on the attention weights and in the heads, models/MMCTransformer.py:41-93).

The masks are counter-based (csrc/dropout.cuh).  Pinned here: the mask function against its numpy restatement, every
kernel that applies a mask against torch arithmetic with THE SAME mask, the attention forward / backward against
torch.autograd over softmax -> mask -> @V, and the whole training step's gradients against autograd over the fp32
restatement of the reference's train graph (oracle.mmct.forward(drop=...)) fed with the step's own masks."""
import ctypes
import math

import numpy as np
import pytest
import torch

import dropout_ref

pytestmark = pytest.mark.gpu
DEV = "cuda"
LOG2E = 1.4426950408889634
P = 0.1
SCALE = 1.0 / (1.0 - P)


def _lib():
    from repurpose_b200 import _lib
    return _lib, _lib.load()


def _drop(L, a, b, p=P):
    return L.RpDropout(a, b, p)


def _mask(L, lib, a, b, shape, p=P):
    n = int(np.prod(shape))
    out = torch.empty(n, dtype=torch.uint8, device=DEV)
    L.check(lib.rp_dropout_mask_u8(ctypes.byref(_drop(L, a, b, p)), n, L.ptr(out), L.cur_stream()), "mask")
    return out.view(*shape).bool()


def _rel(got, ref):
    return ((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("n", [1, 2, 7, 4096, 1_000_001])
def test_mask_matches_the_numpy_restatement(n):
    L, lib = _lib()
    for a, b, p in ((0x12345678, 0x9ABCDEF0, 0.1), (0, 0, 0.1), (0xFFFFFFFF, 1, 0.5), (77, 78, 0.25)):
        got = _mask(L, lib, a, b, (n,), p).cpu().numpy().astype(np.uint8)
        ref = dropout_ref.keep_mask(a, b, p, n)
        assert np.array_equal(got, ref), (n, a, b, p)


def test_mask_statistics_and_stream_separation():
    """keep rate within 5 sigma of 1 - p; streams of different sites / steps are uncorrelated; the attention keep bits are
    the same function of the element index as the byte mask"""
    from repurpose_b200.train import dropout_keys
    L, lib = _lib()
    n = 1 << 22
    masks = []
    for site, step in ((0, 0), (1, 0), (0, 1), (63, 7)):
        a, b = dropout_keys(1234, step, site)
        m = _mask(L, lib, a, b, (n,)).float()
        rate = m.mean().item()
        assert abs(rate - (1 - 6554 / 65536)) < 5 * math.sqrt(P * (1 - P) / n), (site, step, rate)
        # neighbouring elements (the two halves of one hash, and consecutive hashes) are uncorrelated
        for lag in (1, 2, 512, 2048):
            c = ((m[:-lag] - rate) * (m[lag:] - rate)).mean().item() / (rate * (1 - rate))
            assert abs(c) < 5 / math.sqrt(n), (site, step, lag, c)
        masks.append(m)
    for i in range(len(masks)):
        for j in range(i + 1, len(masks)):
            r = masks[i].mean().item()
            c = ((masks[i] - r) * (masks[j] - r)).mean().item() / (r * (1 - r))
            assert abs(c) < 5 / math.sqrt(n), (i, j, c)
    a, b = dropout_keys(1234, 0, 0)
    words = torch.empty(n // 32, dtype=torch.int32, device=DEV)
    L.check(lib.rp_attn_dropout_bits(ctypes.byref(_drop(L, a, b)), n // 32, L.ptr(words), L.cur_stream()), "bits")
    w = words.to(torch.int64) & 0xFFFFFFFF
    bits = ((w[:, None] >> torch.arange(32, device=DEV)) & 1).reshape(-1)
    assert torch.equal(bits.float(), masks[0])


@pytest.mark.parametrize("M,N,K", [(300, 512, 512), (1000, 2048, 512), (2049, 512, 2048), (100, 256, 512), (57632, 256, 256)])
def test_gemm_dropout_epilogues(M, N, K):
    """epilogue 1: dropout(relu(x W^T + b)) as bf16; epilogue 3: resid + dropout(x W^T + b) as fp32 (nn.TransformerEncoderLayer
    _ff_block / _sa_block), with the mask of element row * N + column"""
    L, lib = _lib()
    g = torch.Generator(device=DEV).manual_seed(M + N)
    x = (torch.randn(M, K, device=DEV, generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=DEV, generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g) * 0.1
    resid = torch.randn(M, N, device=DEV, generator=g)
    a, b = 0xDEADBEEF, 0x1234567
    keep = _mask(L, lib, a, b, (M, N))
    drop = _drop(L, a, b)
    y = x.float() @ w.float().t() + bias
    out1 = torch.full((M, N), float("nan"), dtype=torch.bfloat16, device=DEV)
    L.check(lib.rp_gemm_bf16_dropout(1, L.ptr(x), K, L.ptr(w), K, L.ptr(out1), N, L.ptr(bias), 0, 0, M, N, K, ctypes.byref(drop),
                                     L.cur_stream()), "gemm drop 1")
    ref1 = torch.relu(y) * keep * SCALE
    assert _rel(out1, ref1) < 8e-3
    assert ((out1.float() == 0) | keep).all()                      # every dropped element is an exact zero
    assert ((out1.float() != 0) == (keep & (ref1.bfloat16().float() != 0))).float().mean().item() > 0.9999
    out3 = torch.full((M, N), float("nan"), device=DEV)
    L.check(lib.rp_gemm_bf16_dropout(3, L.ptr(x), K, L.ptr(w), K, L.ptr(out3), N, L.ptr(bias), L.ptr(resid), N, M, N, K,
                                     ctypes.byref(drop), L.cur_stream()), "gemm drop 3")
    ref3 = resid + y * keep * SCALE
    assert _rel(out3, ref3) < 2e-5 * max(1.0, math.sqrt(K / 512))
    assert torch.equal(out3[~keep], resid[~keep])                  # dropped: the residual passes through untouched
    # p = 0 through the same entry point = the plain epilogue
    out0 = torch.empty(M, N, device=DEV)
    off = _drop(L, a, b, 0.0)
    L.check(lib.rp_gemm_bf16_dropout(3, L.ptr(x), K, L.ptr(w), K, L.ptr(out0), N, L.ptr(bias), L.ptr(resid), N, M, N, K,
                                     ctypes.byref(off), L.cur_stream()), "gemm drop off")
    plain = torch.empty(M, N, device=DEV)
    L.check(lib.rp_gemm_bf16(3, L.ptr(x), K, L.ptr(w), K, L.ptr(plain), N, L.ptr(bias), L.ptr(resid), N, M, N, K, L.cur_stream()),
            "gemm")
    assert torch.equal(out0, plain)


def test_feature_norm_dropout_and_relu_backward():
    """rp_layernorm512_dropout (mode 2): f = dropout(relu(LN(x))), both head LayerNorms of the dropped row; the ReLU
    backward through the stored dropout(relu(.)) activation scales the surviving gradient by 1 / (1 - p)"""
    L, lib = _lib()
    M = 777
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(M, 512, device=DEV, generator=g) * 2 + 0.3
    gs = [torch.randn(512, device=DEV, generator=g) * 0.2 + 1 for _ in range(3)]
    bs = [torch.randn(512, device=DEV, generator=g) * 0.2 for _ in range(3)]
    a, b = 5, 6
    keep = _mask(L, lib, a, b, (M, 512))
    f = torch.empty(M, 512, device=DEV)
    y1 = torch.empty(M, 512, dtype=torch.bfloat16, device=DEV)
    y2 = torch.empty(M, 512, dtype=torch.bfloat16, device=DEV)
    L.check(lib.rp_layernorm512_dropout(2, L.ptr(x), M, M, L.ptr(gs[0]), L.ptr(bs[0]), L.ptr(gs[1]), L.ptr(bs[1]), L.ptr(gs[2]),
                                        L.ptr(bs[2]), 0, L.ptr(f), L.ptr(y1), L.ptr(y2), ctypes.byref(_drop(L, a, b)),
                                        L.cur_stream()), "ln drop")
    F = torch.nn.functional
    f_ref = torch.relu(F.layer_norm(x, (512,), gs[0], bs[0])) * keep * SCALE
    assert torch.allclose(f, f_ref, atol=2e-5, rtol=1e-5)
    assert _rel(y1, F.layer_norm(f_ref, (512,), gs[1], bs[1])) < 8e-3
    assert _rel(y2, F.layer_norm(f_ref, (512,), gs[2], bs[2])) < 8e-3
    dy = torch.randn(M, 512, device=DEV, generator=g)
    d32 = dy.clone()
    L.check(lib.rp_relu_bwd_scaled(L.ptr(d32), L.ptr(f), M * 512, 1, SCALE, L.cur_stream()), "relu f32")
    assert torch.allclose(d32, torch.where(f > 0, dy * SCALE, torch.zeros_like(dy)), rtol=1e-6, atol=0)
    act = f.bfloat16()
    d16 = dy.bfloat16()
    d16b = d16.clone()
    L.check(lib.rp_relu_bwd_scaled(L.ptr(d16), L.ptr(act), M * 512, 0, SCALE, L.cur_stream()), "relu bf16")
    ref16 = torch.where(act.float() > 0, d16b.float() * SCALE, torch.zeros_like(dy)).bfloat16()
    assert torch.equal(d16, ref16)
    cs = torch.empty(512, device=DEV)
    sc = torch.empty(int(lib.rp_train_scratch_bytes()), dtype=torch.uint8, device=DEV)
    L.check(lib.rp_relu_bwd_colsum_scaled(L.ptr(d16b), L.ptr(act), M, 512, SCALE, L.ptr(cs), L.ptr(sc), sc.numel(),
                                          L.cur_stream()), "relu colsum")
    assert torch.equal(d16b, ref16)
    assert torch.allclose(cs, ref16.float().sum(0), atol=2e-3 * math.sqrt(M), rtol=1e-4)


def test_layernorm_backward_hands_dropout_s_backward_to_the_branch():
    """dh (the residual stream's gradient) is untouched by the mask; dh_bf16 and the bias column sums are dropout's
    backward of it: what out_proj / linear2 receive as dY"""
    L, lib = _lib()
    M = 1234
    g = torch.Generator(device=DEV).manual_seed(9)
    x = torch.randn(M, 512, device=DEV, generator=g)
    dy = torch.randn(M, 512, device=DEV, generator=g)
    gamma = torch.randn(512, device=DEV, generator=g) * 0.2 + 1
    old = torch.randn(M, 512, device=DEV, generator=g)
    a, b = 99, 100
    keep = _mask(L, lib, a, b, (M, 512))
    sc = torch.empty(int(lib.rp_train_scratch_bytes()), dtype=torch.uint8, device=DEV)
    outs = []
    for drop in (None, _drop(L, a, b)):
        dh = old.clone()
        dh16 = torch.empty(M, 512, dtype=torch.bfloat16, device=DEV)
        cs, dg, db = (torch.empty(512, device=DEV) for _ in range(3))
        args = (L.ptr(x), L.ptr(dy), L.ptr(gamma), M, 1e-5, 1, L.ptr(dh), L.ptr(dh16), L.ptr(cs), L.ptr(dg), L.ptr(db), L.ptr(sc),
                sc.numel())
        if drop is None:
            L.check(lib.rp_layernorm512_bwd_acc(*args, L.cur_stream()), "ln bwd")
        else:
            L.check(lib.rp_layernorm512_bwd_acc_dropout(*args, ctypes.byref(drop), L.cur_stream()), "ln bwd drop")
        outs.append((dh, dh16, cs, dg, db))
    (dh0, dh16_0, cs0, dg0, db0), (dh1, dh16_1, cs1, dg1, db1) = outs
    assert torch.equal(dh0, dh1) and torch.equal(dg0, dg1) and torch.equal(db0, db1)
    ref16 = (dh1 * keep * SCALE).bfloat16()
    assert torch.equal(dh16_1, ref16)
    assert torch.allclose(cs1, (dh1 * keep * SCALE).sum(0), atol=1e-3 * math.sqrt(M), rtol=1e-4)
    assert torch.equal(dh16_0, dh0.bfloat16())


def test_head_out_backward_scaled():
    L, lib = _lib()
    M = 999
    g = torch.Generator(device=DEV).manual_seed(4)
    a2 = torch.relu(torch.randn(M, 256, device=DEV, generator=g)).bfloat16()
    w = torch.randn(256, device=DEV, generator=g)
    dlog = torch.randn(M, device=DEV, generator=g)
    sc = torch.empty(int(lib.rp_train_scratch_bytes()), dtype=torch.uint8, device=DEV)
    da2 = torch.empty(M, 256, dtype=torch.bfloat16, device=DEV)
    dw, db = torch.empty(256, device=DEV), torch.empty(1, device=DEV)
    L.check(lib.rp_head_out_bwd_scaled(L.ptr(dlog), L.ptr(a2), L.ptr(w), M, SCALE, L.ptr(da2), L.ptr(dw), L.ptr(db), L.ptr(sc),
                                       sc.numel(), L.cur_stream()), "head bwd")
    ref = torch.where(a2.float() > 0, dlog[:, None] * w[None, :] * SCALE, torch.zeros(M, 256, device=DEV))
    assert _rel(da2, ref) < 8e-3
    assert torch.allclose(dw, (dlog[:, None] * a2.float()).sum(0), atol=1e-3 * math.sqrt(M), rtol=1e-4)
    assert torch.allclose(db, dlog.sum().reshape(1), atol=1e-3, rtol=1e-4)


def _attention_case(B, T, lens, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    H, D = 8, 512
    qkv = torch.randn(B, T, 3 * D, device=DEV, generator=g)
    qkv[..., :D] *= 1.5
    qs = qkv.clone()
    qs[..., :D] *= LOG2E / 8
    d_o = torch.randn(B, T, D, device=DEV, generator=g).bfloat16()
    return H, D, qs.bfloat16(), d_o, torch.tensor(lens, dtype=torch.int32, device=DEV)


@pytest.mark.parametrize("B,T,lens", [(1, 128, [128]), (2, 200, [200, 77]), (2, 333, [333, 129]), (1, 700, [641]),
                                      (2, 1801, [1801, 1211])])
def test_fmha_dropout_forward_and_backward_match_autograd(B, T, lens):
    """out = (keep o softmax(S)) V / (1 - p) — nn.MultiheadAttention's dropout on the attention weights — and dq, dk, dv
    against torch.autograd over exactly that expression with the mask the kernels read"""
    L, lib = _lib()
    H, D, qs, d_o, lens_t = _attention_case(B, T, lens, 31 * T + B)
    ld = ((T + 127) // 128) * 4
    bits = torch.empty(B * H * T, ld, dtype=torch.int32, device=DEV)
    a, b = 0xABCDEF01, 0x10FEDCBA
    L.check(lib.rp_attn_dropout_bits(ctypes.byref(_drop(L, a, b)), bits.numel(), L.ptr(bits), L.cur_stream()), "bits")
    w = bits.view(B, H, T, ld).to(torch.int64) & 0xFFFFFFFF
    keep = ((w[..., None] >> torch.arange(32, device=DEV)) & 1).reshape(B, H, T, ld * 32)[..., :T].bool()
    first = torch.from_numpy(dropout_ref.keep_mask(a, b, P, 8 * ld * 32)).bool().view(8, ld * 32)[:, :T]
    assert torch.equal(keep[0, 0, :8].cpu(), first)                         # the site's first rows against the restatement
    o = torch.empty(B, T, D, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(B, H, T, device=DEV)
    L.check(lib.rp_fmha_train_dropout(L.ptr(qs), L.ptr(qs) + 2 * D, L.ptr(qs) + 4 * D, L.ptr(o), 3 * D, D, B, H, T,
                                      L.ptr(lens_t), L.ptr(lse), L.ptr(bits), ld, P, L.cur_stream()), "fmha drop")
    q = (qs[..., :D].float() * (8 / LOG2E)).view(B, T, H, 64).transpose(1, 2).requires_grad_(True)
    k = qs[..., D:2 * D].float().view(B, T, H, 64).transpose(1, 2).requires_grad_(True)
    v = qs[..., 2 * D:].float().view(B, T, H, 64).transpose(1, 2).requires_grad_(True)
    s = (q @ k.transpose(-1, -2)) / 8
    keymask = torch.arange(T, device=DEV)[None, :] >= lens_t[:, None]
    s = s.masked_fill(keymask[:, None, None, :], float("-inf"))
    p_ref = torch.softmax(s, -1)
    o_ref = ((p_ref * keep * SCALE) @ v).transpose(1, 2).reshape(B, T, D)
    assert _rel(o, o_ref) < 2e-2, _rel(o, o_ref)
    assert torch.allclose(lse, torch.logsumexp(s, -1) * LOG2E, atol=2e-2, rtol=1e-3)   # the row sum ignores the mask
    o_ref.backward(d_o.float())
    dqkv = torch.full((B, T, 3 * D), float("nan"), dtype=torch.bfloat16, device=DEV)
    dsum = torch.empty(B, H, T, device=DEV)
    L.check(lib.rp_fmha_bwd_dropout(L.ptr(qs), L.ptr(qs) + 2 * D, L.ptr(qs) + 4 * D, L.ptr(o), L.ptr(d_o), L.ptr(lse),
                                    L.ptr(dsum), L.ptr(dqkv), L.ptr(dqkv) + 2 * D, L.ptr(dqkv) + 4 * D, 3 * D, D, 3 * D, B, H, T,
                                    L.ptr(lens_t), L.ptr(bits), ld, P, L.cur_stream()), "fmha bwd drop")
    torch.cuda.synchronize()
    assert not torch.isnan(dqkv.float()).any()
    for name, got, ref in (("dq", dqkv[..., :D], q.grad), ("dk", dqkv[..., D:2 * D], k.grad), ("dv", dqkv[..., 2 * D:], v.grad)):
        ref = ref.transpose(1, 2).reshape(B, T, D)
        err = _rel(got, ref)
        assert err < 3e-2, (name, err)
    for bb, n in enumerate(lens):
        assert (dqkv[bb, n:, D:].float() == 0).all()


def _train_setup(layers, lens, seed):
    from oracle import synth
    from repurpose_b200.models.MMCTransformer import MMCTransformer
    torch.manual_seed(seed)
    cfg = dict(synth.MODEL_CFG, self_num_layers=layers)
    model = MMCTransformer(**cfg).to(DEV)
    batch = synth.make_batch(lens, seed=seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    batch["labels"] = (torch.rand(len(lens), max(lens), generator=g) < 0.3).float()
    batch = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in batch.items()}
    return cfg, model, batch


def _reference_with_masks(ts, sd, batch, B, autocast=False):
    """autograd over the fp32 restatement of the reference's TRAIN graph, every nn.Dropout replaced by the mask the
    step's kernels applied (TrainStep.keep_mask / attention_keep_mask)"""
    from oracle import losses as ol
    from oracle import mmct
    Bn, T = batch["visual_feats"].shape[:2]
    shapes = {"drop1": (Bn, T, 512), "ffn": (Bn, T, 2048), "drop2": (Bn, T, 512), "feats": (Bn, T, 512),
              "cls1": (Bn, T, 256), "cls2": (Bn, T, 256), "reg1": (Bn, T, 256), "reg2": (Bn, T, 256)}

    def drop(site, x):
        name, layer = site
        keep = ts.attention_keep_mask(layer) if name == "attn" else ts.keep_mask(ts.site(name, layer), shapes[name])
        return x * keep * ts.drop_scale

    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and not k.endswith(".pe")) for k, v in sd.items()}
    with torch.enable_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        logits, _, _ = mmct.forward.__wrapped__(sd, batch, drop=drop)
        loss = ol.losses(batch["masks"], logits.float(), batch["labels"]) / B
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in sd.items()}


@pytest.mark.parametrize("layers,lens", [(2, [300, 170]), (4, [500, 413])])
def test_train_step_gradients_with_dropout(layers, lens):
    from repurpose_b200.train import TrainStep
    cfg, model, batch = _train_setup(layers, lens, seed=40 + layers)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    B = len(lens)
    ts = TrainStep(model, lr=1e-3, dropout=P, seed=2024)
    loss = ts.loss_and_grads(batch, batch_size=B)
    ref_loss, ref_grads = _reference_with_masks(ts, sd, batch, B)
    _, amp_grads = _reference_with_masks(ts, sd, batch, B, autocast=True)
    assert abs(float(loss) - float(ref_loss)) <= 2e-2 * abs(float(ref_loss)), (float(loss), float(ref_loss))
    rows = []
    for name, p in ts.model.named_parameters():
        if name.startswith("reg_head."):
            continue
        got, ref = ts.grad(name), ref_grads[name]
        l2 = float((got - ref).norm() / ref.norm())
        cos = float((got * ref).sum() / (got.norm() * ref.norm()))
        amp = float((amp_grads[name].float() - ref).norm() / ref.norm())
        rows.append((l2, cos, amp, name))
    rows.sort(reverse=True)
    print("dropout on — largest relative L2 gradient errors (ours, cosine, torch bf16 autocast):",
          [(round(a, 4), round(c, 5), round(m, 4), n) for a, c, m, n in rows[:5]])
    for l2, cos, amp, name in rows:
        assert cos >= 0.99, (name, cos)
        assert l2 <= max(2e-2, 1.3 * amp), (name, l2, amp)
    # the same step index draws the same masks: gradients equal to the rounding of the fused attention backward's dQ sums
    # (fp32 adds in L2, no fixed order) — and bit-identical with the deterministic kernels; the next index draws new masks
    g1 = ts.opt.grad.clone()
    ts.loss_and_grads(batch, batch_size=B)
    assert float((g1 - ts.opt.grad).norm() / g1.norm()) < 1e-3
    from repurpose_b200 import _lib
    lib = _lib.load()
    try:
        _lib.check(lib.rp_set_attn_bwd_deterministic(1), "rp_set_attn_bwd_deterministic")
        ts.loss_and_grads(batch, batch_size=B)
        g1 = ts.opt.grad.clone()
        ts.loss_and_grads(batch, batch_size=B)
        assert torch.equal(g1, ts.opt.grad)
    finally:
        _lib.check(lib.rp_set_attn_bwd_deterministic(-1), "rp_set_attn_bwd_deterministic")
    m0 = ts.keep_mask(ts.site("ffn", 0), (B, max(lens), 2048)).clone()
    ts.step_index += 1
    m1 = ts.keep_mask(ts.site("ffn", 0), (B, max(lens), 2048))
    agree = (m0 == m1).float().mean().item()
    assert abs(agree - (0.9 * 0.9 + 0.1 * 0.1)) < 5e-3, agree
    ts.loss_and_grads(batch, batch_size=B)
    assert not torch.equal(g1, ts.opt.grad)


def test_dropout_changes_the_forward_and_training_still_converges():
    """train mode differs from eval mode, draws change from step to step, and iterations with dropout reduce the loss"""
    from repurpose_b200.train import TrainStep
    cfg, model, batch = _train_setup(2, [256, 200], seed=61)
    ts = TrainStep(model, lr=3e-4, dropout=0.0)
    l_eval = ts.forward(batch)[1].clone()
    ts = TrainStep(model, lr=3e-4, dropout=P, seed=7)
    l0 = ts.forward(batch)[1].clone()
    assert torch.equal(l0, ts.forward(batch)[1])            # same step, same draw
    ts.step_index = 1
    l1 = ts.forward(batch)[1].clone()
    assert not torch.equal(l0, l_eval) and not torch.equal(l0, l1)
    valid = batch["masks"][:, 0, :].bool()
    dev = (l0 - l_eval)[valid].abs().mean().item() / l_eval[valid].abs().mean().item()
    assert 1e-3 < dev < 1.0, dev                             # a perturbation, not garbage
    ts.step_index = 0
    losses = [float(ts.step(batch)) for _ in range(10)]
    print("loss per iteration with dropout:", [round(x, 4) for x in losses])
    assert min(losses[-3:]) < 0.85 * losses[0], losses
