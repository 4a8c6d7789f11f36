#!/usr/bin/env python
"""Multi-GPU check of BASELINE config 3 at small scale: videos with test-split-like lengths are
sharded per video across the ranks (torchrun, one rank per GPU, NCCL), every rank ends with the
full result list after ONE all-gather, and rank 0 compares it with a single-rank run.
    torchrun --nproc-per-node 2 tools/sharded_check.py [--videos 48]
"""
import argparse, os, sys
from pathlib import Path
import torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from repurpose_b200 import synth
from repurpose_b200.models.MMCTransformer import MMCTransformer
from repurpose_b200 import scheduler as S

ap = argparse.ArgumentParser(); ap.add_argument("--videos", type=int, default=48); a = ap.parse_args()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
torch.manual_seed(0)
m = MMCTransformer(512, 2048, 384, 512, 2, 3, 3, 8)
m.load_state_dict(synth.bias_reg_head({k: v.clone() for k, v in m.state_dict().items()}))
m = m.to(f"cuda:{lr}").eval()
lens = synth.sample_lengths(a.videos, seed=1, t_max=600)
g = torch.Generator().manual_seed(2)
videos = [{"visual_feats": torch.randn(t, 512, generator=g), "audio_feats": torch.randn(t, 2048, generator=g),
           "text_feats": torch.randn(t, 384, generator=g), "video_id": i} for i, t in enumerate(lens)]
out = S.run_sharded_inference(m, videos, synth.TEST_CFG, batch_size=8)
assert len(out) == a.videos and [o["video_id"] for o in out] == list(range(a.videos))
if rank == 0:
    # single-rank evaluation of the same videos, one by one (no sharding, no gather)
    bad = 0
    for i, v in enumerate(videos):
        r = m.inference_(S.collate([v]), synth.TEST_CFG, to_host=True)[0]
        if r["labels"].tolist() != out[i]["labels"].tolist():
            bad += 1
    print(f"sharded_check world={world} videos={a.videos} segments={sum(len(o['scores']) for o in out)} mismatching_videos={bad}")
    assert bad == 0
dist.barrier(); dist.destroy_process_group()
