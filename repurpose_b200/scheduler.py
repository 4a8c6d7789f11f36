"""Batch scheduling for the inference path: whole videos are sharded across ranks (one process per
GPU), each rank runs length-bucketed batches through `MMCTransformer.inference_device`, and the
per-rank fixed-slot segment lists are combined with ONE all-gather (NCCL over NVLink on the GPU
box; gloo in the CPU tests).  Nothing here has a counterpart in the reference — its inference is
single-process, batch_size=1 (inference.py:29-31) and its DDP eval never gathers (main.py:696-716).
SURVEY.md §8(e).

Slot layout per video (float32, SLOT = 1 + 4*K values):
    [count, (start, end, score, label) * K]      K = kcap, identical on every rank
"""
from __future__ import annotations

from typing import Iterable, Sequence

import torch
import torch.distributed as dist


def video_cost(T: int) -> float:
    """Algorithmic FLOPs of one video of T valid steps."""
    return 104_989_696.0 * T + 32_768.0 * T * T


MODEL_FLOPS_PER_S = 7.5e14       # what one B200 sustains on this path (bench.py `model_tflops`)
FEATURE_BYTES_PER_STEP = 11776   # fp32 visual 512 + audio 2048 + text 384 per feature step


def shard_videos(lengths: Sequence[int], world_size: int, h2d_gbs: Sequence[float] | None = None,
                 bytes_per_step: int = FEATURE_BYTES_PER_STEP) -> list[list[int]]:
    """Greedy longest-processing-time assignment of whole videos to ranks; deterministic (ties by
    video index).  Returns, per rank, the global video indices it owns (sorted by length, longest
    first, so consecutive videos make low-padding batches).

    h2d_gbs: measured host->device bandwidth of every rank (`measure_h2d_gbs`, all-gathered).  A video then
    costs a rank max(compute time, upload time): on boxes where some GPUs reach the pinned host memory over a
    slower path (profiles/r02_topo_probe.json: 23 vs 36 GB/s per GPU with all eight active) the faster ranks
    take more videos instead of waiting for the slow ones at the all-gather.  Every rank must pass the same list."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0.0] * world_size
    shards: list[list[int]] = [[] for _ in range(world_size)]

    def seconds(i, r):
        t = video_cost(int(lengths[i])) / MODEL_FLOPS_PER_S
        if h2d_gbs is not None:
            t = max(t, int(lengths[i]) * bytes_per_step / (float(h2d_gbs[r]) * 1e9))
        return t

    for i in order:
        r = min(range(world_size), key=lambda k: (load[k] + seconds(i, k), k))
        shards[r].append(i)
        load[r] += seconds(i, r)
    return shards


def measure_h2d_gbs(device, mb: int = 256, reps: int = 4, group=None) -> list[float]:
    """Pinned-host -> device bandwidth of every rank, measured with all ranks copying at the same time (that is
    when shared paths show), all-gathered: the `h2d_gbs` argument of `shard_videos`.  ~50 ms."""
    n = mb * 1024 * 1024 // 4
    src = torch.empty(n, dtype=torch.float32).pin_memory()
    dst = torch.empty(n, dtype=torch.float32, device=device)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(device)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        dist.barrier(group)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(device)
    mine = torch.tensor([n * 4 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9], dtype=torch.float32, device=device)
    if world == 1:
        return [float(mine[0])]
    out = torch.empty(world, dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return [float(x) for x in out.tolist()]


def make_batches(indices: Sequence[int], batch_size: int) -> list[list[int]]:
    """Chunk a (length-sorted) shard into batches of at most batch_size videos."""
    return [list(indices[i:i + batch_size]) for i in range(0, len(indices), batch_size)]


def pack_slots(segments, scores, labels, counts) -> torch.Tensor:
    """[B,K,2], [B,K], [B,K] (int), [B] (int) -> [B, 1+4K] float32 slots (device op, no sync)."""
    B, K = scores.shape
    body = torch.cat([segments.float(), scores.float().unsqueeze(-1),
                      labels.float().unsqueeze(-1)], dim=-1).reshape(B, 4 * K)
    return torch.cat([counts.float().unsqueeze(-1), body], dim=1).contiguous()


def unpack_slots(slots: torch.Tensor) -> list[dict]:
    """Inverse of pack_slots on a host or device tensor -> list of dicts with tensors trimmed to count."""
    slots = slots.cpu()
    out = []
    K = (slots.shape[1] - 1) // 4
    for row in slots:
        k = int(row[0].item())
        body = row[1:].reshape(K, 4)[:k]
        out.append({"segments": body[:, :2].clone(), "scores": body[:, 2].clone(),
                    "labels": body[:, 3].to(torch.int64)})
    return out


def gather_slots(local_slots: torch.Tensor, owned: Sequence[int], shards: Sequence[Sequence[int]],
                 group=None) -> torch.Tensor:
    """One all-gather of equal-sized per-rank slot blocks; returns [n_videos, SLOT] in GLOBAL video
    order on every rank.  `local_slots[i]` belongs to global video `owned[i]`; `shards` is the
    deterministic shard map every rank computed from the same lengths."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n_max = max(len(s) for s in shards)
    slot = local_slots.shape[1]
    block = torch.zeros(n_max, slot, dtype=torch.float32, device=local_slots.device)
    block[:len(owned)] = local_slots
    if world == 1:
        gathered = block.unsqueeze(0)
    else:
        gathered = torch.empty(world, n_max, slot, dtype=torch.float32, device=local_slots.device)
        dist.all_gather_into_tensor(gathered.view(world * n_max, slot), block, group=group)
    n_total = sum(len(s) for s in shards)
    out = torch.zeros(n_total, slot, dtype=torch.float32, device=local_slots.device)
    for r, idxs in enumerate(shards):
        if len(idxs):
            out[torch.as_tensor(list(idxs), device=out.device)] = gathered[r, :len(idxs)]
    return out


def collate(videos: Sequence[dict], pin: bool = False) -> dict:
    """Pad per-video host features to the batch maximum and build the left-aligned mask — the batch
    dict contract of dataset/RepurposeClip.py:450-533 (preprocessing) / :997-1038 (collate_fn_test)."""
    lens = [int(v["visual_feats"].shape[0]) for v in videos]
    T = max(lens)
    batch = {}
    for key in ("visual_feats", "audio_feats", "text_feats"):
        dim = videos[0][key].shape[1]
        x = torch.zeros(len(videos), T, dim, dtype=torch.float32)
        if pin:
            x = x.pin_memory()
        for i, v in enumerate(videos):
            src = torch.as_tensor(v[key], dtype=torch.float32)[:lens[i]]   # a shorter text track stays zero-padded
            x[i, :src.shape[0]] = src
        batch[key] = x
    batch["masks"] = (torch.arange(T)[None, :] < torch.tensor(lens)[:, None])[:, None, :]
    batch["labels"] = torch.zeros(len(videos), T)
    batch["segments"] = torch.zeros(len(videos), T, 2)
    batch["video_id"] = [v.get("video_id", i) for i, v in enumerate(videos)]
    batch["duration"] = lens
    return batch


class InferencePipeline:
    """Multi-buffered (default: three staging slots) host->device->host inference over a stream of collated HOST batches (the
    batched replacement of the reference's `for batch in loader: .to('cuda'); inference_()` loop,
    inference.py:39-47).  The H2D copy of batch i+1 runs on a side stream while batch i computes;
    each batch's fixed-slot result block comes back with one async D2H copy into pinned memory.
    Host tensors should be pinned (`collate(..., pin=True)`) for the copies to overlap."""

    FEATS = ("visual_feats", "audio_feats", "text_feats", "masks")
    RAGGED = ("visual_feats", "audio_feats", "text_feats", "row_offsets", "text_offsets", "text_lens", "lens")

    def __init__(self, model, test_cfg: dict, depth: int = 3):
        self.model, self.cfg, self.depth = model, test_cfg, max(2, depth)
        self.dev = model.device
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self._bufs = [dict() for _ in range(self.depth)]     # device staging, reused per slot
        self._free = [None] * self.depth                      # event: compute that read slot is done
        self._host_out = [None] * self.depth

    def _buffer(self, bufs: dict, key: str, shape, dtype) -> torch.Tensor:
        """grow-only device staging buffer for `key`, viewed with `shape`"""
        n = 1
        for d in shape:
            n *= int(d)
        flat = bufs.get(key)
        if flat is None or flat.dtype != dtype or flat.numel() < n:
            flat = bufs[key] = torch.empty(max(n, 1), dtype=dtype, device=self.dev)
        return flat[:n].view(*shape)

    def _stage(self, slot: int, batch: dict) -> dict:
        """enqueue the H2D copies of `batch` into staging slot `slot` on the copy stream"""
        bufs = self._bufs[slot]
        dbatch = dict(batch)
        with torch.cuda.stream(self.copy_stream):
            if self._free[slot] is not None:
                self.copy_stream.wait_event(self._free[slot])
            parts = batch.get("parts")
            for k in (self.RAGGED if batch.get("ragged") else self.FEATS):
                if parts is not None and k in parts:
                    # zero-copy ragged batch: every video goes straight to its row offset
                    rows = sum(int(t.shape[0]) for t in parts[k])
                    dst = self._buffer(bufs, k, (max(rows, 1), int(parts[k][0].shape[1])), parts[k][0].dtype)
                    pos = 0
                    for t in parts[k]:
                        n = int(t.shape[0])
                        if n:
                            dst[pos:pos + n].copy_(t, non_blocking=True)
                        pos += n
                else:
                    src = batch[k]
                    dst = self._buffer(bufs, k, tuple(src.shape), src.dtype)
                    dst.copy_(src, non_blocking=True)
                dbatch[k] = dst
            dbatch.pop("parts", None)
            ready = torch.cuda.Event()
            ready.record(self.copy_stream)
        return dbatch, ready

    def run(self, batches, raw: bool = False):
        """yields, per input batch and in order, the list of per-video result dicts (CPU tensors);
        with raw=True the batch's fixed-slot block `[B, 1 + 4K]` (a CPU tensor) instead."""
        main = torch.cuda.current_stream(self.dev)
        it = iter(batches)
        pending = []        # (slot, batch, host_slots, done_event)
        staged = []         # (slot, dbatch, ready_event, batch)
        slot = 0

        def stage_next():
            nonlocal slot
            try:
                b = next(it)
            except StopIteration:
                return False
            dbatch, ready = self._stage(slot, b)
            staged.append((slot, dbatch, ready, b))
            slot = (slot + 1) % self.depth
            return True

        for _ in range(self.depth - 1):                    # keep depth-1 uploads queued ahead of the compute
            stage_next()
        while staged:
            s, dbatch, ready, b = staged.pop(0)
            stage_next()                                   # H2D of later batches overlaps this compute
            main.wait_event(ready)
            r = self.model.inference_device(dbatch, self.cfg)
            slots = pack_slots(r["segments"], r["scores"], r["labels"], r["counts"])
            host = self._host_out[s]
            if host is None or host.shape != slots.shape:
                host = self._host_out[s] = torch.empty(slots.shape, dtype=slots.dtype).pin_memory()
            host.copy_(slots, non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            self._free[s] = done
            pending.append((b, host, done))
            if len(pending) >= self.depth:                 # retire the oldest while the GPU works
                yield self._finish(*pending.pop(0), raw)
        while pending:
            yield self._finish(*pending.pop(0), raw)

    @staticmethod
    def _finish(batch, host, done, raw=False):
        done.synchronize()
        if raw:
            return host.clone()
        out = unpack_slots(host)
        for o, vid, dur in zip(out, batch["video_id"], batch["duration"]):
            o["video_id"], o["duration"] = vid, dur
        return out


def run_sharded_inference(model, videos: Sequence[dict], test_cfg: dict, batch_size: int = 32,
                          kcap: int | None = None, group=None, return_slots: bool = False, shards=None):
    """Shard `videos` (list of per-video host dicts with visual_feats/audio_feats/text_feats [T,C])
    across the ranks of `group`, run inference on this rank's share, all-gather, and return one
    result dict per video in the caller's order (identical on every rank).  return_slots=True also
    returns the gathered `[n_videos, 1 + 4K]` slot tensor (device) for `metrics.atiou`.  `shards`: a shard map
    every rank computed identically beforehand (e.g. to decide which feature files to load); entries of
    `videos` this rank does not own are then never touched except for their length."""
    import numpy as np
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lengths = [int(v["visual_feats"].shape[0]) for v in videos]
    if kcap is None:
        kcap = max(1, max(int(np.ceil((l // 60) * test_cfg["max_seg_per_min"])) for l in lengths))
    if shards is None:
        shards = shard_videos(lengths, world)
    owned = shards[rank]
    dev = model.device
    # this rank's share: length-bucketed batches through the pipelined host->device path; every video is
    # copied straight from its own (ideally pinned) arrays to its row offset of the device batch
    from .features import ragged_batch
    local = torch.zeros(len(owned), 1 + 4 * kcap, dtype=torch.float32)
    pos = 0
    pipe = InferencePipeline(model, test_cfg)
    for slots in pipe.run((ragged_batch([videos[i] for i in idxs]) for idxs in make_batches(owned, batch_size)),
                          raw=True):
        local[pos:pos + slots.shape[0], :slots.shape[1]] = slots
        pos += slots.shape[0]
    local = local.to(dev)
    merged = gather_slots(local, owned, shards, group)
    out = unpack_slots(merged)
    for i, o in enumerate(out):
        o["video_id"] = videos[i].get("video_id", i)
        o["duration"] = lengths[i]
    return (out, merged) if return_slots else out
