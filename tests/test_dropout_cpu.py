"""CPU-side pins of the train-mode dropout (no GPU): the host key derivation against the numpy restatement of the mask
function, and the PLACEMENT of every dropout site of oracle.mmct.forward(drop=...) — the graph the GPU gradient-parity
tests differentiate — against the reference's own module in train() mode (oracle/_ref, staged unmodified) with its
nn.Dropout modules replaced by recorded deterministic masks, and against torch's nn.MultiheadAttention for the
attention-weight dropout."""
import math

import numpy as np
import pytest
import torch

import dropout_ref
from oracle import build_ref, mmct, synth


def test_host_keys_and_restatement_agree():
    from repurpose_b200.train import _fmix32, dropout_keys
    xs = np.array([0, 1, 2, 0xFFFFFFFF, 0x9E3779B1, 123456789], dtype=np.uint32)
    assert [int(v) for v in dropout_ref.fmix32(xs.copy())] == [_fmix32(int(v)) for v in xs]
    seen = set()
    for seed in (0, 1, 2 ** 40 + 5):
        for step in range(4):
            for site in range(70):
                seen.add(dropout_keys(seed, step, site))
    assert len(seen) == 3 * 4 * 70                      # every (seed, step, site) owns a stream
    m = dropout_ref.keep_mask(*dropout_keys(3, 0, 0), 0.1, 1 << 20)
    assert abs(m.mean() - (1 - 6554 / 65536)) < 5 * math.sqrt(0.09 / (1 << 20))


def _small_case():
    torch.manual_seed(11)
    cfg = dict(synth.MODEL_CFG, self_num_layers=2)
    batch = synth.make_batch([40, 23], seed=12)
    return cfg, batch


def test_dropout_sites_match_the_reference_module_in_train_mode():
    ref = build_ref.import_reference()
    if ref is None:
        pytest.skip("oracle/_ref not staged")
    cfg, batch = _small_case()
    rmodel = ref.MMCTransformer(**cfg)
    rmodel.train()                                       # main.py:285
    for layer in rmodel.multimodal_encoder.layers:
        layer.self_attn.dropout = 0.0                    # the attention-weight dropout has no nn.Dropout module (next test)
    site_of = {"feature_map.3": ("feats", 0), "cls_head.3": ("cls1", 0), "cls_head.6": ("cls2", 0),
               "reg_head.3": ("reg1", 0), "reg_head.6": ("reg2", 0)}
    for l in range(cfg["self_num_layers"]):
        p = f"multimodal_encoder.layers.{l}."
        site_of[p + "dropout1"], site_of[p + "dropout"], site_of[p + "dropout2"] = ("drop1", l), ("ffn", l), ("drop2", l)
    drops = {n: m for n, m in rmodel.named_modules() if isinstance(m, torch.nn.Dropout)}
    assert set(drops) == set(site_of), sorted(set(drops) ^ set(site_of))   # every nn.Dropout of the reference is a site
    assert all(m.p == 0.1 for m in drops.values())
    masks = {}

    def patched(name):
        def fwd(x):
            g = torch.Generator().manual_seed(1000 + sorted(site_of).index(name))
            keep = torch.rand(x.shape, generator=g) >= 0.1
            masks[site_of[name]] = keep
            return x * keep / 0.9
        return fwd

    for n, m in drops.items():
        m.forward = patched(n)
    with torch.no_grad():
        _, r_logits, r_offsets, _, _, r_feats = rmodel(batch)
    assert len(masks) == len(site_of)
    sd = {k: v.detach() for k, v in rmodel.state_dict().items()}

    def drop(site, x):
        return x if site[0] == "attn" else x * masks[site] / 0.9

    logits, offsets, feats = mmct.forward(sd, batch, drop=drop)
    assert torch.allclose(logits, r_logits, atol=2e-5, rtol=1e-5)
    assert torch.allclose(offsets, r_offsets, atol=2e-5, rtol=1e-5)
    assert torch.allclose(feats, r_feats, atol=2e-5, rtol=1e-5)
    e_logits, _, _ = mmct.forward(sd, batch)
    assert not torch.allclose(logits, e_logits, atol=1e-3)   # the masks did something


def test_attention_weight_dropout_is_softmax_then_mask_then_value_product(monkeypatch):
    """nn.MultiheadAttention(dropout=0.1).train(): torch's explicit path (need_weights=True) drops the softmax OUTPUT
    with F.dropout and multiplies the dropped weights with V; the row sum is taken before the mask.  The oracle's
    ('attn', l) site sits at the same place."""
    torch.manual_seed(5)
    B, T, D, H = 2, 37, 512, 8
    mha = torch.nn.MultiheadAttention(D, H, dropout=0.1, batch_first=True).train()
    x = torch.randn(B, T, D)
    kpm = torch.zeros(B, T, dtype=torch.bool)
    kpm[1, 29:] = True
    rec = {}

    def fake_dropout(inp, p=0.5, training=True, inplace=False):
        g = torch.Generator().manual_seed(77)
        keep = torch.rand(inp.shape, generator=g) >= p
        rec["keep"], rec["p"], rec["rowsum"] = keep, p, inp.sum(-1)
        return inp * keep / (1 - p)

    monkeypatch.setattr(torch.nn.functional, "dropout", fake_dropout)
    with torch.no_grad():
        out, _ = mha(x, x, x, key_padding_mask=kpm, need_weights=True)
    assert rec["p"] == 0.1 and torch.allclose(rec["rowsum"], torch.ones_like(rec["rowsum"]), atol=1e-5)
    keep = rec["keep"].view(B, H, T, T)
    # the same through the oracle's encoder-layer arithmetic (attention part only)
    sd = {"w": mha.in_proj_weight.detach(), "b": mha.in_proj_bias.detach()}
    qkv = torch.nn.functional.linear(x, sd["w"], sd["b"])
    q, k, v = (t.view(B, T, H, 64).transpose(1, 2) for t in qkv.split(D, dim=-1))
    s = (q @ k.transpose(-2, -1)) / 8
    s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
    o = (torch.softmax(s, -1) * keep / 0.9) @ v
    o = o.transpose(1, 2).reshape(B, T, D)
    mine = torch.nn.functional.linear(o, mha.out_proj.weight.detach(), mha.out_proj.bias.detach())
    assert torch.allclose(mine, out, atol=2e-5, rtol=1e-5)
