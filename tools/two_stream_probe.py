#!/usr/bin/env python
"""Probe: does running the batch as TWO half batches on two streams (the kernels of one half filling the idle SMs and
tails of the other's: the LayerNorm-fused GEMMs only co-schedule 33 clusters of 4 = 132 of 148 SMs) beat one launch
sequence over the whole batch?  Device-resident inputs, forward + decode + Soft-NMS, CUDA events.

    python tools/two_stream_probe.py [--B 32] [--T 1801] [--iters 10] [--parts 2]
"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from repurpose_b200 import synth  # noqa: E402
from repurpose_b200.models.MMCTransformer import MMCTransformer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--T", type=int, default=1801)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--parts", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    sd = None
    models = []
    for _ in range(a.parts + 1):
        m = MMCTransformer(**synth.MODEL_CFG)
        if sd is None:
            sd = synth.bias_reg_head({k: v.clone() for k, v in m.state_dict().items()})
        m.load_state_dict(sd)
        models.append(m.to(dev).eval())
    cfg = synth.TEST_CFG
    host = synth.make_batch([a.T] * a.B, seed=1000)
    full = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in host.items()}
    n = a.B // a.parts
    parts = [{k: (v[i * n:(i + 1) * n].contiguous() if torch.is_tensor(v) else v[i * n:(i + 1) * n]) for k, v in full.items()}
             for i in range(a.parts)]
    streams = [torch.cuda.Stream(dev) for _ in range(a.parts)]

    def one():
        return models[0].inference_device(full, cfg)

    def split():
        cur = torch.cuda.current_stream()
        outs = []
        for s, m, p in zip(streams, models[1:], parts):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                outs.append(m.inference_device(p, cfg))
        for s in streams:
            cur.wait_stream(s)
        return outs

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters

    out = {"B": a.B, "T": a.T, "parts": a.parts}
    for rep in range(2):
        out[f"one_sequence_ms_{rep}"] = timeit(one)
        out[f"split_streams_ms_{rep}"] = timeit(split)
    r1, r2 = one(), split()
    torch.cuda.synchronize()
    c1 = r1["counts"].cpu()
    c2 = torch.cat([r["counts"].cpu() for r in r2])
    out["same_counts"] = bool(torch.equal(c1, c2))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
