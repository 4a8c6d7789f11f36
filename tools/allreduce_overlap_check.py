#!/usr/bin/env python
"""Two or more ranks under torchrun: the bucketed, overlapped gradient all-reduce of TrainStep against the single flat
all-reduce after the backward pass — same averaged gradients (bit-identical with the deterministic attention backward: a
per-element sum over the ranks does not depend on how the buffer is cut), and the step time of both.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/allreduce_overlap_check.py
"""
import copy
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from repurpose_b200 import synth  # noqa: E402
from repurpose_b200.models.MMCTransformer import MMCTransformer  # noqa: E402
from repurpose_b200.train import TrainStep  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, T = 16, 1801
torch.manual_seed(0)
base = MMCTransformer(**synth.MODEL_CFG).to(dev)
batch = synth.make_batch([T] * B, seed=100 + rank)
g = torch.Generator().manual_seed(7 + rank)
batch["labels"] = (torch.rand(B, T, generator=g) < 0.3).float()
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
out = {"world": world, "B": B, "T": T}
grads = {}
for mode in (False, True):
    ts = TrainStep(copy.deepcopy(base), lr=1e-4, weight_decay=1e-4, dropout=0.1, seed=1000 + rank, deterministic=True)
    ts.overlap_allreduce = mode
    ts.step(batch, batch_size=B)
    grads[mode] = ts.opt.grad.clone()
    for _ in range(2):
        ts.step(batch, batch_size=B)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ts.step(batch, batch_size=B)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["overlapped_ms" if mode else "flat_ms"] = float(t.item())
    del ts
out["bit_identical"] = bool(torch.equal(grads[False], grads[True]))
out["max_abs_diff"] = float((grads[False] - grads[True]).abs().max())
# every rank holds the same averaged gradient
ref = grads[True].clone()
dist.broadcast(ref, src=0)
out["same_on_all_ranks"] = bool(torch.equal(ref, grads[True]))
if rank == 0:
    print(json.dumps(out))
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/allreduce_overlap.json").write_text(json.dumps(out) + "\n")
dist.destroy_process_group()
