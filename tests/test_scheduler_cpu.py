"""CPU suite, part 3: host-side scheduling logic — per-video sharding, batch bucketing, slot
packing, and the single all-gather, exercised with world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import synth
from repurpose_b200 import scheduler as S


def test_shard_is_a_balanced_partition():
    lens = synth.sample_lengths(1000, seed=3)
    for world in (1, 2, 4, 8):
        shards = S.shard_videos(lens, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(len(lens)))
        loads = [sum(S.video_cost(lens[i]) for i in s) for s in shards]
        assert max(loads) / (sum(loads) / world) < 1.02
        assert S.shard_videos(lens, world) == shards  # deterministic
        for s in shards:  # longest first inside a shard -> low padding per batch
            assert [lens[i] for i in s] == sorted((lens[i] for i in s), reverse=True)


def test_sample_lengths_follow_test_split_distribution():
    lens = synth.sample_lengths(20000, seed=0)
    assert min(lens) >= 28 and max(lens) == 1801
    assert 1150 < sum(lens) / len(lens) < 1280  # reference test split: mean 1211


def test_make_batches_and_collate():
    assert S.make_batches(list(range(7)), 3) == [[0, 1, 2], [3, 4, 5], [6]]
    vids = [{"visual_feats": torch.ones(t, 4), "audio_feats": torch.ones(t, 6),
             "text_feats": torch.ones(t, 2), "video_id": f"v{t}"} for t in (5, 3)]
    b = S.collate(vids)
    assert b["visual_feats"].shape == (2, 5, 4) and b["masks"].shape == (2, 1, 5)
    assert b["masks"][1, 0].tolist() == [True, True, True, False, False]
    assert b["visual_feats"][1, 3:].abs().sum() == 0 and b["duration"] == [5, 3]


def test_slot_roundtrip():
    B, K = 3, 4
    seg = torch.rand(B, K, 2)
    sc = torch.rand(B, K)
    lab = torch.randint(0, 1801, (B, K), dtype=torch.int32)
    cnt = torch.tensor([4, 0, 2], dtype=torch.int32)
    out = S.unpack_slots(S.pack_slots(seg, sc, lab, cnt))
    for b in range(B):
        k = int(cnt[b])
        assert torch.equal(out[b]["segments"], seg[b, :k]) and torch.equal(out[b]["scores"], sc[b, :k])
        assert out[b]["labels"].tolist() == lab[b, :k].tolist() and out[b]["labels"].dtype == torch.int64


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_videos, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lens = synth.sample_lengths(n_videos, seed=7)
        shards = S.shard_videos(lens, world)
        owned = shards[rank]
        K = 3
        # slot content is a pure function of the GLOBAL video index, so every rank can check all rows
        def slot_of(i):
            return torch.tensor([float(i % (K + 1))] + [float(i * 10 + j) for j in range(4 * K)])
        local = torch.stack([slot_of(i) for i in owned]) if owned else torch.zeros(0, 1 + 4 * K)
        merged = S.gather_slots(local, owned, shards)
        expect = torch.stack([slot_of(i) for i in range(n_videos)])
        q.put((rank, bool(torch.equal(merged, expect)), len(owned)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_videos", [1, 5, 64])   # 1: one rank owns nothing
def test_all_gather_reconstructs_global_order_world2(n_videos):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_videos, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) == n_videos


def test_bandwidth_weighted_sharding_and_device_picking():
    """Boxes whose GPUs reach the pinned host memory at different speeds (profiles/r02_topo_probe.json): upload-bound
    videos go preferentially to the fast ranks, every video is owned exactly once, and a job on fewer GPUs than
    the box has spreads over both halves of the host."""
    import numpy as np
    from repurpose_b200.affinity import pick_device
    from repurpose_b200.scheduler import shard_videos
    lens = [int(x) for x in np.random.default_rng(0).integers(100, 1801, 999)]
    speeds = [23.0] * 4 + [36.0] * 4
    plain, weighted = shard_videos(lens, 8), shard_videos(lens, 8, speeds)
    for shards in (plain, weighted):
        assert sorted(i for sh in shards for i in sh) == list(range(len(lens)))
    steps = lambda sh: sum(lens[i] for i in sh)
    slow, fast = steps(weighted[0]), steps(weighted[7])
    assert 1.35 < fast / slow < 1.75                        # ~36 / 23
    assert max(map(steps, plain)) / min(map(steps, plain)) < 1.05
    assert shard_videos(lens, 8, speeds) == weighted        # deterministic: every rank builds the same map
    buses = [0x1b, 0x40, 0x53, 0x66, 0x9c, 0xc0, 0xd1, 0xe5]
    assert [pick_device(r, 4, buses) for r in range(4)] == [0, 1, 4, 5]
    assert [pick_device(r, 2, buses) for r in range(2)] == [0, 4]
    assert [pick_device(r, 8, buses) for r in range(8)] == list(range(8))
    assert pick_device(0, 1, buses) == 0 and pick_device(3, 4, []) == 3
