#!/usr/bin/env python
"""BASELINE.json configs[4] (training-step variant): forward + backward + Adam of the Repurpose.yaml model on synthetic
features, one process per GPU, gradients averaged with one flat all-reduce.

    python tools/train_bench.py [--B 16] [--T 1801] [--steps 5]
    python -m torch.distributed.run --nproc-per-node N tools/train_bench.py ...
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def run_train_bench(B, T, steps, warmup=2, rank=0, world=1, dev=None, dropout=0.1):
    from repurpose_b200 import synth
    from repurpose_b200.models.MMCTransformer import MMCTransformer
    from repurpose_b200.train import TrainStep
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(0)
    model = MMCTransformer(**synth.MODEL_CFG).to(dev)
    ts = TrainStep(model, lr=1e-4, weight_decay=1e-4, dropout=dropout, seed=1000 + rank)
    batch = synth.make_batch([T] * B, seed=100 + rank)
    g = torch.Generator().manual_seed(7 + rank)
    batch["labels"] = (torch.rand(B, T, generator=g) < 0.3).float()
    batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    losses = []
    for _ in range(warmup):
        losses.append(float(ts.step(batch, batch_size=B)))
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    fwd_ms = 0.0
    e0.record()
    for _ in range(steps):
        loss = ts.step(batch, batch_size=B)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    losses.append(float(loss))
    # forward alone, for the split
    e0.record()
    for _ in range(steps):
        ts.forward(batch)
    e2.record()
    torch.cuda.synchronize()
    fwd_ms = e0.elapsed_time(e2) / steps
    ts.profile_begin()
    ts.step(batch, batch_size=B)
    prof = {k: {"ms": round(v[0], 3), "launches": v[1]} for k, v in sorted(ts.profile_end().items(), key=lambda kv: -kv[1][0])}
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    flops = 3.0 * B * (104_989_696.0 * T + 32_768.0 * T * T)   # forward + 2x for the backward (SURVEY 8d forward count)
    return {"batch_per_gpu": B, "seq_len": T, "n_gpus": world, "ms_per_step": ms, "forward_ms": fwd_ms,
            "videos_per_s": world * B / (ms * 1e-3), "model_tflops_per_gpu": flops / (ms * 1e-3) / 1e12,
            "loss_first_last": [losses[0], losses[-1]], "kernel_classes_ms": prof,
            "activation_gb": 16 * B * T * 14.3e3 / 1e9,
            "dropout": dropout,
            "what": "forward (activations kept) + masked focal loss + backward with the flat gradient buffer all-reduced in buckets underneath + Adam; "
                    + ("train mode: nn.Dropout(%.2f) at every site of the reference graph (counter-based masks)" % dropout
                       if dropout > 0 else "dropout off (the eval-mode graph)")}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--T", type=int, default=1801)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--dropout", type=float, default=0.1)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    r = run_train_bench(a.B, a.T, a.steps, rank=rank, world=world, dropout=a.dropout)
    if rank == 0:
        print(json.dumps(r))
    if world > 1:
        dist.destroy_process_group()
