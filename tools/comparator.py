#!/usr/bin/env python
"""On-box comparator: the REFERENCE GRAPH executed by PyTorch itself
on the same B200 — the module's own `nn.TransformerEncoder` / `nn.Sequential` parameter store called
through torch in eval / no_grad mode (torch's fused encoder fast path -> its SDPA / cuDNN kernels), fp32
as shipped and with weights + inputs cast to bf16 — followed by the reference-style per-video decode
loop (models/MMCTransformer.py:181-275: sigmoid, threshold, sort, top-k, Soft-NMS on the host), against
this repo's path on the same batch.  Also times torch's scaled_dot_product_attention backends at the
attention shape of the bench (B=32, H=8, T=1801, d=64, bf16): that is the library kernel our FMHA has
to beat.

    python tools/comparator.py [--B 32] [--T 1801] [--iters 5] [--out gpurun_out/comparator.json]

Test / measurement infrastructure: it imports oracle/ for the host-side decode loop.
"""
import argparse
import copy
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import mmct  # noqa: E402
from oracle.softnms import soft_nms_intervals_oracle  # noqa: E402
from repurpose_b200 import synth  # noqa: E402
from repurpose_b200.models.MMCTransformer import MMCTransformer  # noqa: E402

DEV = "cuda:0"


@torch.no_grad()
def eager_forward(model, batch):
    """models/MMCTransformer.py:109-151 op for op, executed by torch on the module's own sub-modules."""
    x = torch.cat([batch["visual_feats"], batch["audio_feats"], batch["text_feats"]], dim=-1)
    x = model.input_projection(x)
    x = model.input_norm(x)
    x = x + model.positional_encoding.pe[:, :x.size(1)].to(x.dtype)
    pad = ~batch["masks"].squeeze(1)
    x = model.multimodal_encoder(x, src_key_padding_mask=pad)
    x = model.encoder_norm(x)
    feats = model.feature_map(x)
    return model.cls_head(feats), model.reg_head(feats), feats


@torch.no_grad()
def eager_inference(model, batch, cfg):
    """reference inference_ (:231-275): forward, then a Python loop with one host sync per video."""
    logits, offsets, _ = eager_forward(model, batch)
    out = []
    for i, dur in enumerate(batch["duration"]):
        r = mmct.decode_single_video(batch["masks"][i], logits[i].float(), offsets[i].float(), cfg)
        max_seg = int(np.ceil((int(dur) // 60) * cfg["max_seg_per_min"]))
        keep = soft_nms_intervals_oracle(r["scores"].cpu().numpy(), r["segments"].cpu().numpy(),
                                         cfg.get("nms_sigma", 0.5), cfg.get("min_score", 0.001), max_seg)
        out.append(keep)
    return out


def time_ms(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (time.perf_counter() - t0) * 1e3 / iters


def sdpa_table(B, H, T, iters):
    """torch SDPA backends on the bench's attention shape, dense (no mask) like our all-full batch."""
    from torch.nn.attention import SDPBackend, sdpa_kernel
    q = torch.randn(B, H, T, 64, device=DEV, dtype=torch.bfloat16)
    k, v = torch.randn_like(q), torch.randn_like(q)
    flops = 4.0 * B * H * T * T * 64
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    out = {}
    for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION),
                     ("efficient", SDPBackend.EFFICIENT_ATTENTION)):
        try:
            with sdpa_kernel([be]):
                for _ in range(3):
                    F.scaled_dot_product_attention(q, k, v)
                torch.cuda.synchronize()
                tot = 0.0
                for _ in range(iters):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    F.scaled_dot_product_attention(q, k, v)
                    e1.record()
                    torch.cuda.synchronize()
                    tot += e0.elapsed_time(e1)
            ms = tot / iters
            out[name] = {"ms": ms, "tflops": flops / ms / 1e9}
        except Exception as e:  # backend not available for this shape / build
            out[name] = {"error": str(e).splitlines()[0][:200]}
    return out


def ours_fmha(B, H, T, iters):
    from repurpose_b200 import _lib
    from repurpose_b200._lib import check, cur_stream, ptr
    lib = _lib.load()
    D = H * 64
    qkv = torch.randn(B, T, 3 * D, device=DEV)
    qkv[..., :D] *= 1.4426950408889634 / 8
    qkv = qkv.bfloat16()
    o = torch.empty(B, T, D, dtype=torch.bfloat16, device=DEV)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

    def run():
        check(lib.rp_fmha(ptr(qkv), ptr(qkv) + D * 2, ptr(qkv) + 2 * D * 2, ptr(o), 3 * D, 3 * D, 3 * D, D,
                          T * 3 * D, T * 3 * D, T * 3 * D, T * D, B, H, T, T, 0, 0, 0, 0, 0, cur_stream()), "fmha")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / iters
    return {"ms": ms, "tflops": 4.0 * B * H * T * T * 64 / ms / 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--T", type=int, default=synth.MAX_SEQ_LEN)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "comparator.json"))
    ap.add_argument("--skip-model", action="store_true")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    res = {"B": args.B, "T": args.T, "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "tf32_matmul": bool(torch.backends.cuda.matmul.allow_tf32)}
    res["sdpa_bf16_dense"] = sdpa_table(args.B, 8, args.T, args.iters)
    res["sdpa_bf16_dense"]["ours_rp_fmha"] = ours_fmha(args.B, 8, args.T, args.iters)
    if not args.skip_model:
        torch.manual_seed(0)
        model = MMCTransformer(**synth.MODEL_CFG)
        model.load_state_dict(synth.bias_reg_head({k: v.clone() for k, v in model.state_dict().items()}))
        model = model.to(DEV).eval()
        batch = synth.make_batch([args.T] * args.B, seed=1, T=args.T)
        dbatch = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in batch.items()}
        cfg = synth.TEST_CFG
        # ours
        ms, _ = time_ms(lambda: model(dbatch), args.iters)
        res["ours_forward_ms"] = ms
        _, wall = time_ms(lambda: model.inference_(dbatch, cfg, to_host=True), args.iters)
        res["ours_inference_wall_ms"] = wall
        # the reference graph through torch, fp32 as shipped
        ms, _ = time_ms(lambda: eager_forward(model, dbatch), args.iters)
        res["eager_fp32_forward_ms"] = ms
        _, wall = time_ms(lambda: eager_inference(model, dbatch, cfg), max(1, args.iters // 2), warm=1)
        res["eager_fp32_inference_wall_ms"] = wall
        # parity of the two on the same batch (valid rows: all)
        lo, off, _ = eager_forward(model, dbatch)
        _, lg, og, _, _, _ = model(dbatch)
        res["max_abs_logit_diff_vs_eager_fp32"] = float((lg - lo).abs().max())
        res["max_abs_offset_diff_vs_eager_fp32"] = float((og - off).abs().max())
        # bf16 weights + inputs (library tensor-core path)
        m16 = copy.deepcopy({k: v for k, v in model.state_dict().items()})
        model16 = MMCTransformer(**synth.MODEL_CFG)
        model16.load_state_dict(m16)
        model16 = model16.to(DEV).eval().bfloat16()
        b16 = {k: (v.bfloat16() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in dbatch.items()}
        ms, _ = time_ms(lambda: eager_forward(model16, b16), args.iters)
        res["eager_bf16_forward_ms"] = ms
        _, wall = time_ms(lambda: eager_inference(model16, b16, cfg), max(1, args.iters // 2), warm=1)
        res["eager_bf16_inference_wall_ms"] = wall
        res["videos_per_s"] = {"ours_inference": args.B / res["ours_inference_wall_ms"] * 1e3,
                               "eager_fp32_inference": args.B / res["eager_fp32_inference_wall_ms"] * 1e3,
                               "eager_bf16_inference": args.B / res["eager_bf16_inference_wall_ms"] * 1e3}
        res["speedup_forward"] = {"vs_eager_fp32": res["eager_fp32_forward_ms"] / res["ours_forward_ms"],
                                  "vs_eager_bf16": res["eager_bf16_forward_ms"] / res["ours_forward_ms"]}
    print(json.dumps(res, indent=1))
    Path(args.out).parent.mkdir(exist_ok=True)
    Path(args.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
