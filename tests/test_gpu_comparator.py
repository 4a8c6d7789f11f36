"""GPU suite: full-depth, full-length parity against the REFERENCE GRAPH executed by PyTorch itself on the
same GPU — the module's own nn.TransformerEncoder / nn.Sequential parameter store called through torch in
eval / no_grad mode (torch's fused encoder fast path and its SDPA kernels; models/MMCTransformer.py:41-55,
109-151) — and the on-box comparator timing SURVEY.md §8d asks for (fp32 as shipped, and bf16 weights +
inputs).  tools/comparator.py runs the same comparison at the bench shape (B=32) with the per-video decode
loop and the SDPA backends; its JSON is what profiles/r02_comparator.json holds.  Timing is reported, not
asserted beyond "the CUDA path is faster"."""
import json
import os
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))
from comparator import eager_forward  # noqa: E402
from oracle import synth  # noqa: E402
from repurpose_b200.models.MMCTransformer import MMCTransformer  # noqa: E402
from test_gpu_model import _elementwise_gate  # noqa: E402

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

    assert not torch.backends.cuda.matmul.allow_tf32          # the comparator must be true fp32
    ref_logits, ref_offsets, ref_feats = eager_forward(m, dbatch)
    _, logits, offsets, _, _, feats = m(dbatch)
    valid = dbatch["masks"][:, 0, :]
    for name, got, ref in (("logits", logits, ref_logits), ("offsets", offsets, ref_offsets), ("feats", feats, ref_feats)):
        g, r = got[valid].float(), ref[valid].float()
        err = (g - r).abs().max().item() / max(r.abs().max().item(), 1e-6)
        assert err < 2e-2, (name, err)                          # north_star: 2e-2 relative (bf16 path vs fp32)
        _elementwise_gate(f"16 layers x T=1801 {name} vs torch eager fp32", g, r)

    ms_ours = _time(lambda: m(dbatch))
    ms_fp32 = _time(lambda: eager_forward(m, dbatch))
    m16 = MMCTransformer(**synth.MODEL_CFG)
    m16.load_state_dict({k: v.cpu() for k, v in m.state_dict().items()})
    m16 = m16.to(DEV).eval().bfloat16()
    b16 = {k: (v.bfloat16() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in dbatch.items()}
    ms_bf16 = _time(lambda: eager_forward(m16, b16))
    out = {"batch": len(lens), "T": T, "lens": lens, "ms_forward_ours": ms_ours, "ms_forward_eager_fp32": ms_fp32,
           "ms_forward_eager_bf16": ms_bf16, "speedup_vs_eager_fp32": ms_fp32 / ms_ours,
           "speedup_vs_eager_bf16": ms_bf16 / ms_ours,
           "note": "forward only; eager = the module's own torch sub-modules (nn.TransformerEncoder fast path -> SDPA) "
                   "on the same GPU"}
    print("\ncomparator:", json.dumps(out))
    d = Path(os.environ.get("GRAFT_REPO_ROOT", ROOT)) / "gpurun_out"
    d.mkdir(exist_ok=True)
    (d / "comparator_test.json").write_text(json.dumps(out))
    assert ms_ours < ms_fp32
