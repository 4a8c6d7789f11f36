"""Feature files and GPU collate (SURVEY.md §8 f2 — the step right before the hot path).

The reference keeps three `.npy` files per video (`[T, C]` float32/float64: CLIP 512-d, PANNs 2048-d,
MiniLM 384-d), slices them by `timeRange`, truncates to a common length, pads every batch to its longest
video on the host and copies the padded fp32 batch to the GPU (dataset/RepurposeClip.py:401-446,
450-533, 962-1038).  Here the host only concatenates the videos' rows back to back (no padding bytes over
PCIe) and one kernel builds the padded bf16 concat `[B, T, 2944]` the input projection consumes
(`rp_forward_ragged`)."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch


def load_video_features(visual_path, audio_path, text_path, time_range=None, n_labels=None,
                        dtype: str = "fp32", pin: bool = False) -> dict:
    """One video's features with the reference's slicing rules (RepurposeClip.__getitem__ :962-994):
    rows [int(t0), int(t1)) of each file when timeRange[0] != 0; duration = min(visual, audio[, labels])
    rows — the text file is NOT part of the minimum and may end up shorter (its missing rows are zeros
    after collation).  dtype="bf16" returns torch bf16 tensors (see `to_bf16`); pin=True returns pinned torch
    tensors so that `ragged_batch` + `InferencePipeline` can DMA them without a host copy."""
    vis = np.load(visual_path, allow_pickle=True)
    aud = np.load(audio_path, allow_pickle=True)
    txt = np.load(text_path, allow_pickle=True)
    if time_range is not None and time_range[0] != 0:
        t0, t1 = int(time_range[0]), int(time_range[1])
        vis, aud, txt = vis[t0:t1, :], aud[t0:t1, :], txt[t0:t1, :]
    n = min(vis.shape[0], aud.shape[0])
    if n_labels is not None:
        n = min(n, int(n_labels))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    video = {"visual_feats": f32(vis[:n]), "audio_feats": f32(aud[:n]), "text_feats": f32(txt[:n]),
             "duration": n}
    if dtype == "bf16":   # rows as the device would round them anyway: half the upload, identical results
        return to_bf16(video, pin=pin)
    if dtype != "fp32":
        raise ValueError(f"dtype must be 'fp32' or 'bf16', got {dtype!r}")
    if pin:
        for k in ("visual_feats", "audio_feats", "text_feats"):
            video[k] = torch.from_numpy(video[k]).pin_memory()
    return video


def collate_ragged(videos: Sequence[dict], pin: bool = True) -> dict:
    """Host side of the GPU collate: back-to-back rows + per-video offsets (no padding).  `videos` are
    dicts with visual_feats [T,Cv], audio_feats [T,Ca], text_feats [Tt,Ct] (numpy or torch, Tt <= T
    allowed) and optional video_id."""
    lens = [int(v["visual_feats"].shape[0]) for v in videos]
    tlens = [min(int(v["text_feats"].shape[0]), l) for v, l in zip(videos, lens)]

    def cat(key, ls):
        dim = int(videos[0][key].shape[1])
        out = torch.empty(max(1, sum(ls)), dim, dtype=torch.float32)
        if pin:
            out = out.pin_memory()
        pos = 0
        for v, l in zip(videos, ls):
            if l:
                out[pos:pos + l] = torch.as_tensor(v[key][:l], dtype=torch.float32)
            pos += l
        return out

    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    toff = np.concatenate([[0], np.cumsum(tlens)[:-1]]).astype(np.int32)
    return {"ragged": True,
            "visual_feats": cat("visual_feats", lens), "audio_feats": cat("audio_feats", lens),
            "text_feats": cat("text_feats", tlens),
            "row_offsets": torch.from_numpy(off), "text_offsets": torch.from_numpy(toff),
            "text_lens": torch.tensor(tlens, dtype=torch.int32),
            "lens": torch.tensor(lens, dtype=torch.int32),
            "video_id": [v.get("video_id", i) for i, v in enumerate(videos)],
            "duration": lens, "labels": None, "segments": None}


def ragged_batch(videos: Sequence[dict]) -> dict:
    """Zero-copy variant of `collate_ragged`: the batch only references the per-video arrays
    (`parts`), and the host->device stage copies each one straight to its row offset of the device
    buffers (`InferencePipeline`), so the host never touches the feature bytes.  Sources should be
    pinned torch tensors for the copies to be asynchronous."""
    lens = [int(v["visual_feats"].shape[0]) for v in videos]
    tlens = [min(int(v["text_feats"].shape[0]), l) for v, l in zip(videos, lens)]
    off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int32)
    toff = np.concatenate([[0], np.cumsum(tlens)[:-1]]).astype(np.int32)
    def as_t(a):  # fp32 like the reference's loader, or bf16 rows converted once with `to_bf16`
        return a if torch.is_tensor(a) and a.dtype == torch.bfloat16 else torch.as_tensor(a, dtype=torch.float32)
    return {"ragged": True,
            "parts": {"visual_feats": [as_t(v["visual_feats"][:l]) for v, l in zip(videos, lens)],
                      "audio_feats": [as_t(v["audio_feats"][:l]) for v, l in zip(videos, lens)],
                      "text_feats": [as_t(v["text_feats"][:l]) for v, l in zip(videos, tlens)]},
            "row_offsets": torch.from_numpy(off), "text_offsets": torch.from_numpy(toff),
            "text_lens": torch.tensor(tlens, dtype=torch.int32), "lens": torch.tensor(lens, dtype=torch.int32),
            "video_id": [v.get("video_id", i) for i, v in enumerate(videos)],
            "duration": lens, "labels": None, "segments": None}


def to_bf16(video: dict, pin: bool = True) -> dict:
    """One-off conversion of a video's feature arrays to bf16 (round to nearest even — exactly what the
    device does to fp32 features before the input projection), e.g. when caching feature files.
    Batches built from such videos with `ragged_batch` take the `rp_forward_ragged_bf16` entry point."""
    out = dict(video)
    for k in ("visual_feats", "audio_feats", "text_feats"):
        t = torch.as_tensor(video[k], dtype=torch.float32).to(torch.bfloat16).contiguous()
        out[k] = t.pin_memory() if pin and torch.cuda.is_available() else t
    return out
