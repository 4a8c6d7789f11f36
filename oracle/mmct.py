"""ORACLE (test infrastructure, see oracle/__init__.py): fp32 CPU restatement of the reference hot
path as plain tensor arithmetic over a reference-schema state dict — no nn.Module, no reference
import.  Each function cites the reference lines it restates (paths relative to
YosubShin/Repurpose).  Pinned against the reference itself by oracle/make_golden.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from .softnms import soft_nms_intervals_oracle

EPS = 1e-5  # nn.LayerNorm default


def positional_table(d_model: int, length: int) -> torch.Tensor:
    """models/MMCTransformer.py:11-17 (sin on even, cos on odd channels, fp32); pure function of
    the position, so `length` may exceed the reference's 5000-row buffer."""
    pe = torch.zeros(length, d_model)
    position = torch.arange(0, length, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def _ln(x, sd, prefix):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], EPS)


def _lin(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _no_drop(site, x):
    return x


def encoder_layer(h, kpm, sd, p, num_heads, drop=_no_drop, layer=0):
    """nn.TransformerEncoderLayer, norm_first=True, ReLU, as configured at models/MMCTransformer.py:41-49.
    kpm [B,T] bool, True = padded key -> -inf before softmax.  Eval mode by default; `drop(site, x)` stands for the
    module's nn.Dropout(0.1) instances under train(): ('attn', l) on the softmax output inside
    nn.MultiheadAttention, ('drop1', l) = dropout1 after out_proj, ('ffn', l) = dropout after the activation,
    ('drop2', l) = dropout2 after linear2 (torch/nn/modules/transformer.py _sa_block / _ff_block)."""
    B, T, D = h.shape
    dk = D // num_heads
    u = _ln(h, sd, p + "norm1")
    qkv = F.linear(u, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"])
    q, k, v = qkv.split(D, dim=-1)
    q = q.view(B, T, num_heads, dk).transpose(1, 2)
    k = k.view(B, T, num_heads, dk).transpose(1, 2)
    v = v.view(B, T, num_heads, dk).transpose(1, 2)
    s = (q @ k.transpose(-2, -1)) / math.sqrt(dk)
    s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
    o = drop(("attn", layer), torch.softmax(s, dim=-1)) @ v
    o = o.transpose(1, 2).reshape(B, T, D)
    h = h + drop(("drop1", layer), _lin(o, sd, p + "self_attn.out_proj"))
    u = _ln(h, sd, p + "norm2")
    h = h + drop(("drop2", layer), _lin(drop(("ffn", layer), torch.relu(_lin(u, sd, p + "linear1"))), sd, p + "linear2"))
    return h


@torch.no_grad()
def forward(sd, batch, num_heads=8, query_chunk=None, drop=_no_drop):
    """MMCTransformer.forward, models/MMCTransformer.py:109-151.
    Returns (cls_logits [B,T,1], offsets [B,T,2], feats [B,T,D]).  `drop(site, x)`: the train-mode nn.Dropout at
    `site` (see encoder_layer; heads: ('feats', 0) = feature_map[3], ('cls1' | 'cls2' | 'reg1' | 'reg2', 0) =
    cls_head[3] / [6], reg_head[3] / [6], :63-93); the default is eval mode."""
    x = torch.cat([batch["visual_feats"], batch["audio_feats"], batch["text_feats"]], dim=-1).float()
    B, T, _ = x.shape
    h = _ln(_lin(x, sd, "input_projection"), sd, "input_norm")               # :121, :124
    D = h.shape[-1]
    pe = sd.get("positional_encoding.pe")
    pe = positional_table(D, T) if pe is None or pe.shape[1] < T else pe[0, :T]
    h = h + pe[None]                                                            # :127
    kpm = (batch["masks"] == 0).squeeze(1)                                      # :132
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("multimodal_encoder.layers."))
    for l in range(n_layers):                                                   # :135-138
        h = encoder_layer(h, kpm, sd, f"multimodal_encoder.layers.{l}.", num_heads, drop, l)
    z = _ln(h, sd, "encoder_norm")                                              # :141
    feats = drop(("feats", 0), torch.relu(_ln(_lin(z, sd, "feature_map.0"), sd, "feature_map.1")))  # :63-68, :144

    def head(name, tag, final_relu):
        a = _ln(feats, sd, name + ".0")
        a = drop((tag + "1", 0), torch.relu(_lin(a, sd, name + ".1")))
        a = drop((tag + "2", 0), torch.relu(_lin(a, sd, name + ".4")))
        a = _lin(a, sd, name + ".7")
        return torch.relu(a) if final_relu else a

    return head("cls_head", "cls", False), head("reg_head", "reg", True), feats  # :71-93, :147-149


@torch.no_grad()
def decode_single_video(mask, logits, offsets, cfg):
    """inference_single_video, models/MMCTransformer.py:181-229.  mask [1,T] bool, logits [T],
    offsets [T,2].  torch.sort is unstable for equal keys; ties are resolved by ascending time step
    here (and in the CUDA kernel) — the reference leaves that order unspecified."""
    prob = (torch.sigmoid(logits).squeeze() * mask).flatten()
    keep = prob > cfg["pre_nms_thresh"]
    idx = keep.nonzero(as_tuple=True)[0]
    p = prob[keep]
    order = torch.argsort(p, descending=True, stable=True)
    k = min(int(cfg["pre_nms_topk"]), idx.numel())
    order = order[:k]
    p, t = p[order], idx[order]
    off = offsets[t]
    left = t - off[:, 0]
    right = t + off[:, 1]
    dur = right - left
    ok = (dur > cfg["duration_thresh"]) & (dur < cfg["duration_thresh_max"])
    return {"segments": torch.stack((left, right), -1)[ok], "scores": p[ok], "labels": t[ok]}


@torch.no_grad()
def inference(sd, batch, cfg, num_heads=8):
    """inference_, models/MMCTransformer.py:231-275, with the reference's CUDA semantics for the
    returned scores (original probability at the kept index).  Also returns the
    decayed Soft-NMS score per kept row (`dscores`) for the 1e-6 check."""
    logits, offsets, _ = forward(sd, batch, num_heads)
    out = []
    for i, (vid, vlen) in enumerate(zip(batch["video_id"], batch["duration"])):
        max_seg = int(np.ceil((int(vlen) // 60) * cfg["max_seg_per_min"]))
        r = decode_single_video(batch["masks"][i], logits[i, :, 0], offsets[i], cfg)
        keep, dsc = soft_nms_intervals_oracle(r["scores"].numpy().copy(), r["segments"].numpy().copy(),
                                              sigma=cfg["nms_sigma"], thresh=cfg["min_score"],
                                              max_seg_num=max_seg, return_scores=True)
        kt = torch.from_numpy(keep)
        out.append({"segments": r["segments"][kt], "scores": r["scores"][kt], "labels": r["labels"][kt],
                    "dscores": torch.from_numpy(dsc), "video_id": vid, "duration": vlen,
                    "ncand": int(r["scores"].numel())})
    return out


def mha_forward(sd, q, k, v, mask, num_heads):
    """MultiHeadAttention.forward, models/transformer.py:52-81 (masked_fill(mask == 0, -1e9))."""
    B, Tq, D = q.shape
    dk = D // num_heads
    qp = _lin(q, sd, "q_linear").view(B, -1, num_heads, dk).transpose(1, 2)
    kp = _lin(k, sd, "k_linear").view(B, -1, num_heads, dk).transpose(1, 2)
    vp = _lin(v, sd, "v_linear").view(B, -1, num_heads, dk).transpose(1, 2)
    s = (qp @ kp.transpose(-2, -1)) / math.sqrt(dk)
    if mask is not None:
        s = s.masked_fill(mask.unsqueeze(1) == 0, -1e9)
    x = torch.softmax(s, dim=-1) @ vp
    x = x.transpose(1, 2).contiguous().view(B, -1, D)
    return _lin(x, sd, "out")


def calculate_tiou(reference_segments, predicted_segments, tiou_thresholds=(0.5,)):
    """utils/metrics.py:82-111: precision of predictions whose best IoU with any GT >= threshold."""
    def iou(a, b):
        inter = max(0, min(a[1], b[1]) - max(a[0], b[0]))
        union = (a[1] - a[0]) + (b[1] - b[0]) - inter
        return inter / union if union != 0 else 0
    best = [max([iou(p, r) for r in reference_segments], default=0) for p in predicted_segments]
    n = len(predicted_segments)
    return {t: (sum(s >= t for s in best) / n if n > 0 else 0) for t in tiou_thresholds}


def atiou(gt_lists, pred_lists, thresholds=(0.5, 0.6, 0.7, 0.8, 0.9)):
    """inference.py:45-55: per-threshold precision averaged over videos, then over thresholds."""
    per = [calculate_tiou(g, p, thresholds) for g, p in zip(gt_lists, pred_lists)]
    by_t = {t: sum(d[t] for d in per) / len(per) for t in thresholds}
    return sum(by_t.values()) / len(by_t), by_t
