#!/usr/bin/env python
"""Latency of the drop-in at the reference's own call pattern (inference.py:31, 39-47): batch size 1, features
already on the device, `model.inference_(batch, cfg)` and the segments brought to the host with .tolist() —
this repo with CUDA-graph replay (default for small batches), with plain launches, and the reference graph
through PyTorch eager (fp32 as shipped, and bf16) with the reference's per-video decode loop.

    python tools/latency_bs1.py [--iters 30] [--out gpurun_out/latency_bs1.json]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from comparator import eager_inference  # noqa: E402
from repurpose_b200 import synth  # noqa: E402
from repurpose_b200.models.MMCTransformer import MMCTransformer  # noqa: E402

DEV = "cuda:0"


def wall_ms(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(iters):
        t0 = time.perf_counter()
        fn()
        t.append((time.perf_counter() - t0) * 1e3)
    t.sort()
    return {"median_ms": t[len(t) // 2], "min_ms": t[0], "p90_ms": t[int(len(t) * 0.9)]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "latency_bs1.json"))
    a = ap.parse_args()
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    model = MMCTransformer(**synth.MODEL_CFG)
    model.load_state_dict(synth.bias_reg_head({k: v.clone() for k, v in model.state_dict().items()}))
    model = model.to(DEV).eval()
    m16 = MMCTransformer(**synth.MODEL_CFG)
    m16.load_state_dict({k: v.cpu() for k, v in model.state_dict().items()})
    m16 = m16.to(DEV).eval().bfloat16()
    cfg = synth.TEST_CFG
    res = {"gpu": torch.cuda.get_device_name(0), "what": "wall clock per call of inference_ at batch 1, device inputs, "
           "segments to the host (.tolist()), median of %d" % a.iters, "cases": []}
    for B, T in ((1, 600), (1, 1200), (1, 1801), (4, 1801)):
        batch = synth.make_batch([T] * B, seed=5, T=T)
        db = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in batch.items()}
        b16 = {k: (v.bfloat16() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in db.items()}

        def ours():
            return [r["segments"].tolist() for r in model.inference_(db, cfg, to_host=True)]
        model.cuda_graphs = "auto"
        g = wall_ms(ours, a.iters)
        seg_graph = ours()
        model.cuda_graphs = False
        e = wall_ms(ours, a.iters)
        seg_plain = ours()
        model.cuda_graphs = "auto"
        case = {"B": B, "T": T, "ours_cuda_graph": g, "ours_plain_launches": e, "same_segments": seg_graph == seg_plain,
                "eager_fp32": wall_ms(lambda: eager_inference(model, db, cfg), max(3, a.iters // 5), warm=1),
                "eager_bf16": wall_ms(lambda: eager_inference(m16, b16, cfg), max(3, a.iters // 5), warm=1)}
        case["speedup_vs_eager_fp32"] = case["eager_fp32"]["median_ms"] / g["median_ms"]
        res["cases"].append(case)
        print(json.dumps(case), flush=True)
    Path(a.out).parent.mkdir(exist_ok=True)
    Path(a.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
