#!/usr/bin/env python
"""On-box comparator for the TRAINING step (BASELINE configs[4]): the reference graph trained by
PyTorch itself on the same B200 — the module's own `nn.TransformerEncoder` / `nn.Sequential` sub-modules in train()
mode (dropout 0.1), the reference's masked focal loss / batch_size (main.py:326), `loss.backward()`,
`torch.optim.Adam.step()` (main.py:190) — in fp32 as the reference runs it and under torch's bf16 autocast, against
`repurpose_b200.train.TrainStep` on the same batch.  Also times torch's cuDNN SDPA forward + backward at the attention
shape of the step: the library kernel pair our attention backward has to be compared with.

    python tools/train_comparator.py [--B 16] [--T 1801] [--iters 3] [--out gpurun_out/train_comparator.json]

Measurement infrastructure: imports oracle/ for the loss restatement.
"""
import argparse
import json
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import losses as ol  # noqa: E402
from repurpose_b200 import synth  # noqa: E402
from repurpose_b200.models.MMCTransformer import MMCTransformer  # noqa: E402

DEV = "cuda:0"


def eager_train_forward(model, batch):
    """models/MMCTransformer.py:109-151 on the module's own sub-modules (train mode: every nn.Dropout active)."""
    x = torch.cat([batch["visual_feats"], batch["audio_feats"], batch["text_feats"]], dim=-1)
    x = model.input_projection(x)
    x = model.input_norm(x)
    x = x + model.positional_encoding.pe[:, :x.size(1)].to(x.dtype)
    x = model.multimodal_encoder(x, src_key_padding_mask=~batch["masks"].squeeze(1))
    x = model.encoder_norm(x)
    feats = model.feature_map(x)
    return model.cls_head(feats), model.reg_head(feats)


def time_ms(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--T", type=int, default=1801)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    B, T = a.B, a.T
    res = {"B": B, "T": T, "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}

    # ---- attention kernels alone: cuDNN SDPA forward + backward vs ours (dense, no padding)
    from torch.nn.attention import SDPBackend, sdpa_kernel
    q, k, v = (torch.randn(B, 8, T, 64, device=DEV, dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
    flops_fwd = 4.0 * B * 8 * T * T * 64
    try:
        with sdpa_kernel([SDPBackend.CUDNN_ATTENTION]):
            o = F.scaled_dot_product_attention(q, k, v)
            do = torch.randn_like(o)
            fwd = time_ms(lambda: F.scaled_dot_product_attention(q, k, v), 10)

            def bwd():
                q.grad = k.grad = v.grad = None
                o.backward(do, retain_graph=True)
            bw = time_ms(bwd, 10)
        res["sdpa_cudnn"] = {"fwd_ms": fwd, "bwd_ms": bw, "fwd_tflops": flops_fwd / fwd / 1e9,
                             "bwd_tflops_algorithmic": 2.5 * flops_fwd / bw / 1e9}
    except Exception as e:  # noqa: BLE001
        res["sdpa_cudnn"] = {"error": repr(e)[:300]}
    del q, k, v

    # ---- the step
    torch.manual_seed(0)
    model = MMCTransformer(**synth.MODEL_CFG).to(DEV)
    batch = synth.make_batch([T] * B, seed=100)
    g = torch.Generator().manual_seed(7)
    batch["labels"] = (torch.rand(B, T, generator=g) < 0.3).float()
    batch = {kk: (vv.to(DEV) if torch.is_tensor(vv) else vv) for kk, vv in batch.items()}

    from repurpose_b200.train import TrainStep
    import copy
    ts = TrainStep(copy.deepcopy(model), lr=1e-4, weight_decay=1e-4, dropout=0.1, seed=1000)
    res["ours_train_step_ms"] = time_ms(lambda: ts.step(batch, batch_size=B), a.iters)
    del ts
    torch.cuda.empty_cache()

    for name, amp in (("eager_bf16_autocast", True), ("eager_fp32", False)):
        m = copy.deepcopy(model).train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-4, weight_decay=1e-4)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                logits, _ = eager_train_forward(m, batch)
            loss = ol.losses(batch["masks"], logits.float(), batch["labels"]) / B
            loss.backward()
            opt.step()
            return loss
        try:
            res[name + "_train_step_ms"] = time_ms(step, a.iters, warmup=1)
            res[name + "_peak_gb"] = torch.cuda.max_memory_allocated() / 1e9
        except Exception as e:  # noqa: BLE001 (out of memory at large B is an answer too)
            res[name + "_train_step_ms"] = None
            res[name + "_error"] = repr(e)[:300]
        del m, opt
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    o = res["ours_train_step_ms"]
    res["speedup"] = {kk: (res[kk + "_train_step_ms"] / o if res.get(kk + "_train_step_ms") else None)
                      for kk in ("eager_bf16_autocast", "eager_fp32")}
    s = json.dumps(res)
    print(s)
    if a.out:
        Path(a.out).write_text(s + "\n")


if __name__ == "__main__":
    main()
