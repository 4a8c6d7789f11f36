#!/usr/bin/env python
"""Diagnostic: per-parameter gradient error of TrainStep vs fp32 autograd of the reference graph, next to the error of
torch's own bf16 autocast over the same graph (the noise floor of any bf16 tensor-core backward)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from oracle import mmct  # noqa: E402
from oracle import losses as ol  # noqa: E402
from test_gpu_train_step import _setup  # noqa: E402


def ref_grads(sd, batch, B, autocast):
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and not k.endswith(".pe")) for k, v in sd.items()}
    with torch.enable_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        logits, _, _ = mmct.forward.__wrapped__(sd, batch)
        loss = ol.losses(batch["masks"], logits.float(), batch["labels"]) / B
    loss.backward()
    return float(loss), {k: v.grad for k, v in sd.items()}


def main():
    layers = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    lens = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [300, 170]
    from repurpose_b200.train import TrainStep
    cfg, model, batch = _setup(layers, lens, seed=5 + layers)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    B = len(lens)
    ts = TrainStep(model, lr=1e-3, dropout=0.0)
    loss = float(ts.loss_and_grads(batch, batch_size=B))
    l32, g32 = ref_grads(sd, batch, B, False)
    l16, g16 = ref_grads(sd, batch, B, True)
    print(f"loss ours {loss:.5f} fp32 {l32:.5f} autocast {l16:.5f}")
    print(f"{'parameter':58s} {'ours max':>9s} {'ours L2':>8s} {'cos':>8s} | {'amp max':>8s} {'amp L2':>8s}")
    for name, p in model.named_parameters():
        if name.startswith("reg_head."):
            continue
        r = g32[name]
        o = ts.grad(name)
        a = g16[name].float()
        mx = lambda x: float((x - r).abs().max() / r.abs().max())
        l2 = lambda x: float((x - r).norm() / r.norm())
        cos = float((o * r).sum() / (o.norm() * r.norm()))
        if ("layers" not in name) or ".0." in name[:32] or f".{layers - 1}." in name[:34]:
            print(f"{name:58s} {mx(o):9.4f} {l2(o):8.4f} {cos:8.5f} | {mx(a):8.4f} {l2(a):8.4f}")


if __name__ == "__main__":
    main()
