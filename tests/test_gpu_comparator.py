"""GPU suite: full-depth, full-length parity against the fp32 restatement of the reference running as
PyTorch eager ON THE SAME GPU (the CPU oracle needs minutes at this size), and the on-box comparator
timing SURVEY.md §8d asks for (reference algorithm, eager fp32 / bf16, same batch, same run).
The timing is reported (gpurun_out/comparator.json), not asserted beyond "the CUDA path is faster"."""
import json
import os
from pathlib import Path

import pytest
import torch

from oracle import mmct, synth
from repurpose_b200.models.MMCTransformer import MMCTransformer

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _time(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def test_full_depth_parity_and_eager_comparator():
    torch.manual_seed(0)
    m = MMCTransformer(**synth.MODEL_CFG).to(DEV).eval()
    T = synth.MAX_SEQ_LEN
    lens = [T, 1500, T, 700]
    batch = synth.make_batch(lens, seed=21, T=T)
    dbatch = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in batch.items()}
    sd = {k: v.detach().to(DEV) for k, v in m.state_dict().items()}

    assert not torch.backends.cuda.matmul.allow_tf32          # the comparator must be true fp32
    ref_logits, ref_offsets, ref_feats = mmct.forward(sd, dbatch)
    _, logits, offsets, _, _, feats = m(dbatch)
    valid = dbatch["masks"][:, 0, :]
    for name, got, ref in (("logits", logits, ref_logits), ("offsets", offsets, ref_offsets), ("feats", feats, ref_feats)):
        g, r = got[valid].float(), ref[valid].float()
        err = (g - r).abs().max().item() / max(r.abs().max().item(), 1e-6)
        assert err < 2e-2, (name, err)                          # north_star: 2e-2 relative (bf16 path vs fp32)

    ms_ours = _time(lambda: m(dbatch))
    ms_fp32 = _time(lambda: mmct.forward(sd, dbatch))
    sd16 = {k: (v.bfloat16() if v.is_floating_point() else v) for k, v in sd.items()}
    b16 = {k: (v.bfloat16() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in dbatch.items()}

    def eager_bf16():
        x = dict(b16)
        # the restatement concatenates and calls .float(): keep the library bf16 path by patching the cat
        x["visual_feats"], x["audio_feats"], x["text_feats"] = b16["visual_feats"], b16["audio_feats"], b16["text_feats"]
        return _forward_bf16(sd16, x)

    ms_bf16 = _time(eager_bf16)
    out = {"batch": len(lens), "T": T, "lens": lens, "ms_forward_ours": ms_ours, "ms_forward_eager_fp32": ms_fp32,
           "ms_forward_eager_bf16": ms_bf16, "speedup_vs_eager_fp32": ms_fp32 / ms_ours,
           "speedup_vs_eager_bf16": ms_bf16 / ms_ours,
           "note": "forward only; eager = oracle/mmct.py restatement of the reference on the same GPU (library GEMMs, "
                   "materialised attention scores)"}
    print("\ncomparator:", json.dumps(out))
    d = Path(os.environ.get("GRAFT_REPO_ROOT", Path(__file__).resolve().parent.parent)) / "gpurun_out"
    d.mkdir(exist_ok=True)
    (d / "comparator.json").write_text(json.dumps(out))
    assert ms_ours < ms_fp32


@torch.no_grad()
def _forward_bf16(sd, batch, num_heads=8):
    """the same op sequence as oracle.mmct.forward without its fp32 upcast (library bf16 tensor-core path)"""
    x = torch.cat([batch["visual_feats"], batch["audio_feats"], batch["text_feats"]], dim=-1)
    B, T, _ = x.shape
    h = mmct._ln(mmct._lin(x, sd, "input_projection"), sd, "input_norm")
    h = h + sd["positional_encoding.pe"][0, :T][None]
    kpm = (batch["masks"] == 0).squeeze(1)
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("multimodal_encoder.layers."))
    for l in range(n_layers):
        h = mmct.encoder_layer(h, kpm, sd, f"multimodal_encoder.layers.{l}.", num_heads)
    z = mmct._ln(h, sd, "encoder_norm")
    return torch.relu(mmct._ln(mmct._lin(z, sd, "feature_map.0"), sd, "feature_map.1"))
