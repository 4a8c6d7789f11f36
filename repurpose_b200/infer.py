"""Batched, sharded replacement of the reference's `inference.py`: same configuration
file, same checkpoint, same printed result, but whole videos are sharded over the ranks (torchrun, one
rank per GPU), length-bucketed batches stream through the pinned zero-copy host->device pipeline, the
per-rank segment lists meet in ONE all-gather and AtIoU is computed on the device.

    python -m repurpose_b200.infer --config_path configs/Repurpose.yaml --resume best.pth
    torchrun --nproc-per-node 8 -m repurpose_b200.infer --config_path ... --resume ... [--bf16-features]

Reference flow: inference.py:22-55 (batch size 1, 24 loader workers, one host sync per video);
dataset rules: dataset/RepurposeClip.py:579-606 (test set, labels from timeRangeOffset), :962-994 (slicing).
"""
from __future__ import annotations

import argparse
import json
import os
from concurrent.futures import ThreadPoolExecutor

import torch
import torch.distributed as dist


def read_test_set(ds_cfg: dict) -> list[dict]:
    """One entry per usable video of `test_dataset.label_path`: feature paths, timeRange, the number of
    label steps the reference derives from timeRangeOffset (int(t1 - t0) + 1,
    dataset/RepurposeClip.py:878-879) and the ground-truth segments (`segmentsOffset`).  Videos without all
    three feature files are skipped, like the reference's availability filter (:595-597)."""
    with open(ds_cfg["label_path"]) as f:
        labels = json.load(f)
    out = []
    for k in labels:
        vid = k["youtube_id"]
        paths = [os.path.join(ds_cfg[key], f"{vid}.npy") for key in ("video_path", "audio_path", "text_path")]
        if not all(os.path.exists(p) for p in paths):
            continue
        tro = k["timeRangeOffset"]
        out.append({"video_id": vid, "paths": paths, "time_range": k["timeRange"],
                    "n_labels": int(tro[1] - tro[0]) + 1, "gt_segments": k["segmentsOffset"]})
    return out


def load_videos(entries, dtype="fp32", workers=16):
    from .features import load_video_features

    def one(e):
        v = load_video_features(*e["paths"], time_range=e["time_range"], n_labels=e["n_labels"], dtype=dtype,
                                pin=torch.cuda.is_available())
        v["video_id"] = e["video_id"]
        return v
    with ThreadPoolExecutor(max_workers=workers) as ex:
        return list(ex.map(one, entries))


def main(argv=None):
    import yaml
    ap = argparse.ArgumentParser()
    ap.add_argument("--config_path", required=True)
    ap.add_argument("--resume", required=True, help="checkpoint with a 'model' state dict (inference.py:33-34)")
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--bf16-features", action="store_true", help="convert feature rows to bf16 at load (identical results)")
    args = ap.parse_args(argv)
    with open(args.config_path) as f:
        cfg = yaml.load(f, Loader=yaml.FullLoader)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    from .affinity import bind_to_gpu_numa, pick_device
    local_rank = pick_device(int(os.environ.get("LOCAL_RANK", "0")), world)
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank = dist.get_rank() if world > 1 else 0

    from .metrics import THRESHOLDS, atiou
    from .models.MMCTransformer import MMCTransformer
    from .scheduler import run_sharded_inference, shard_videos
    model = MMCTransformer(**cfg["model"]).to(f"cuda:{local_rank}")
    model.load_state_dict(torch.load(args.resume, map_location="cpu")["model"])
    model.eval()

    entries = read_test_set(cfg["test_dataset"])
    # every rank only loads the feature files of the videos it owns; lengths come from the label file
    # (upper bound: the reference truncates to the shortest of visual / audio / labels)
    shards = shard_videos([e["n_labels"] for e in entries], world)
    mine = set(shards[rank])
    loaded = load_videos([entries[i] for i in sorted(mine)], "bf16" if args.bf16_features else "fp32")
    by_index = dict(zip(sorted(mine), loaded))
    lengths = torch.zeros(len(entries), dtype=torch.int64)
    for i, v in by_index.items():
        lengths[i] = v["duration"]
    if world > 1:  # everyone needs the true lengths to build the same shard map
        lengths = lengths.cuda()
        dist.all_reduce(lengths)
        lengths = lengths.cpu()
    # placeholder entries for the videos other ranks own (only their length is used on this rank)
    videos = [by_index.get(i) or {"visual_feats": torch.empty(int(lengths[i]), 0), "audio_feats": torch.empty(0, 0),
                                  "text_feats": torch.empty(0, 0), "video_id": entries[i]["video_id"]}
              for i in range(len(entries))]
    _, slots = run_sharded_inference(model, videos, cfg["test_cfg"], batch_size=args.batch_size, return_slots=True,
                                     shards=shards)
    average, per_thr, _ = atiou(slots, [e["gt_segments"] for e in entries], THRESHOLDS)
    if rank == 0:
        print(per_thr)                       # inference.py:53-54
        print(f"average tIoU: {average}")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return average


if __name__ == "__main__":
    main()
