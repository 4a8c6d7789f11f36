#!/usr/bin/env python
"""ORACLE (test infrastructure): pin the three remaining restatements against the REFERENCE ITSELF.

Runs in the build container only (imports the unmodified reference from /root/reference):

  * oracle.mmct.mha_forward            vs  models/transformer.py:37-81  MultiHeadAttention.forward
  * oracle.mmct.calculate_tiou / atiou vs  utils/metrics.py:82-111 and the averaging of inference.py:45-55
  * repurpose_b200.infer.read_test_set + features.load_video_features (host loader)
                                        vs  dataset/RepurposeClip.py:578-606, 962-994 RepurposeClipTest
                                            + collate_fn_test :997-1038 (batch size 1, like inference.py)

and writes tests/golden/{mha,tiou,loader}_cases.npz with the reference's outputs.

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_aux.py
"""
import json
import logging
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"
THRESHOLDS = [0.5, 0.6, 0.7, 0.8, 0.9]


# ---------------------------------------------------------------------------- shared case builders
def mha_weights(seed=31, d_model=512):
    """Seeded parameters of the reference MultiHeadAttention (regenerated, not stored: 4 MB)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name in ("q_linear", "k_linear", "v_linear", "out"):
        sd[f"{name}.weight"] = torch.randn(d_model, d_model, generator=g) * 0.06
        sd[f"{name}.bias"] = torch.randn(d_model, generator=g) * 0.1
    return sd


def mha_cases(seed=32, d_model=512):
    """(name, q, k, v, mask) — self / padding / band / cross with an arbitrary mask / a fully masked row."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    out = []
    x = r(2, 96, d_model)
    out.append(("self_nomask", x, x, x, None))
    x = r(2, 150, d_model)
    lens = torch.tensor([150, 47])
    out.append(("self_padding", x, x, x, (torch.arange(150)[None, :] < lens[:, None])[:, None, :]))
    x = r(1, 200, d_model)
    i = torch.arange(200)
    out.append(("self_band16", x, x, x, ((i[:, None] - i[None, :]).abs() <= 16)[None]))
    q, kv = r(2, 70, d_model), r(2, 133, d_model)
    out.append(("cross_random", q, kv, kv, torch.rand(2, 70, 133, generator=g) > 0.4))
    x = r(1, 64, d_model)
    m = torch.ones(1, 64, 64, dtype=torch.bool)
    m[:, 5] = False
    out.append(("fully_masked_row", x, x, x, m))
    return out


def tiou_cases(seed=41, n=200):
    """(gt list, pred list) pairs in the shapes inference.py feeds calculate_tiou (lists of [s, e])."""
    rng = np.random.default_rng(seed)
    cases = []
    for c in range(n):
        T = float(rng.integers(60, 1801))
        ng = int(rng.integers(0, 7)) if c % 9 else 0
        npred = int(rng.integers(0, 10)) if c % 7 else 0
        gt = [sorted(rng.uniform(0, T, 2).tolist()) for _ in range(ng)]
        pred = []
        for _ in range(npred):
            if gt and rng.random() < 0.5:   # near a ground-truth segment: IoUs land around the thresholds
                s, e = gt[int(rng.integers(0, len(gt)))]
                pred.append([float(np.float32(s + rng.normal(0, 4))), float(np.float32(e + rng.normal(0, 4)))])
            else:
                c0, ln = rng.uniform(0, T), rng.uniform(10, 90)
                pred.append([float(np.float32(c0 - ln / 2)), float(np.float32(c0 + ln / 2))])
        if c == 3:
            gt, pred = [[5.0, 5.0]], [[5.0, 5.0]]          # zero union -> IoU 0 (utils/metrics.py:97)
        if c == 4:
            gt, pred = [[10.0, 20.0]], [[10.0, 20.0]]        # IoU exactly 1
        cases.append((gt, pred))
    return cases


def write_loader_dataset(root: Path, seed=51):
    """A tiny dataset in the reference's on-disk format that exercises every slicing rule:
    timeRange[0] != 0, feature files of different lengths, labels shorter than the features,
    a text track shorter than the others, float64 files, a video with a missing modality."""
    rng = np.random.default_rng(seed)
    dims = {"video_path": 12, "audio_path": 20, "text_path": 8}
    dirs = {k: root / k for k in dims}
    for d in dirs.values():
        d.mkdir(parents=True, exist_ok=True)
    specs = [  # name, rows(vis, aud, txt), timeRange, timeRangeOffset, dtype
        ("plain", (90, 90, 90), [0, 89.0], [0, 89.0], np.float32),
        ("ragged_files", (100, 97, 99), [0, 99.0], [0, 99.0], np.float32),
        ("short_labels", (120, 120, 120), [0, 79.0], [0, 79.0], np.float32),
        ("sliced", (300, 300, 300), [40.0, 171.5], [0, 131.5], np.float32),
        ("sliced_short_text", (300, 300, 110), [60.0, 200.0], [0, 140.0], np.float32),
        ("f64", (75, 80, 75), [0, 74.0], [0, 74.0], np.float64),
        ("no_audio", (50, 0, 50), [0, 49.0], [0, 49.0], np.float32),
    ]
    labels = []
    for name, rows, tr, tro, dt in specs:
        for (k, d), n in zip(dirs.items(), rows):
            if n:
                np.save(d / f"{name}.npy", rng.normal(size=(n, dims[k])).astype(dt))
        gt = [sorted(rng.uniform(0, tro[1], 2).tolist()) for _ in range(3)]
        labels.append({"youtube_id": name, "timeRange": tr, "timeRangeOffset": tro, "segments": gt,
                       "segmentsOffset": gt})
    (root / "test.json").write_text(json.dumps(labels))
    return {"label_path": str(root / "test.json"), **{k: str(v) for k, v in dirs.items()}}


def checksum(a) -> float:
    a = np.asarray(a, dtype=np.float64)
    return float((a * np.cos(np.arange(a.size, dtype=np.float64).reshape(a.shape) * 0.37)).sum())


# ---------------------------------------------------------------------------- the pin itself
def main(ref_dir="/root/reference"):
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref_dir)
    from models.transformer import MultiHeadAttention as RefMHA            # the reference
    from utils.metrics import calculate_tiou as ref_tiou                   # the reference
    from dataset.RepurposeClip import RepurposeClipTest, collate_fn_test   # the reference
    sys.path.insert(0, str(ROOT))
    from oracle import mmct
    from repurpose_b200.features import load_video_features
    from repurpose_b200.infer import read_test_set

    logging.disable(logging.CRITICAL)
    report = []

    # ---- MultiHeadAttention
    sd = mha_weights()
    ref = RefMHA(512, 8).eval()
    missing = ref.load_state_dict(sd, strict=False)
    assert set(missing.missing_keys) <= {"scale"} and not missing.unexpected_keys
    gold = {"weight_checksum": np.array(sum(checksum(v.numpy()) for v in sd.values()))}
    for name, q, k, v, mask in mha_cases():
        with torch.no_grad():
            y = ref(q, k, v, mask)
        o = mmct.mha_forward(sd, q, k, v, mask, 8)
        d = float((y - o).abs().max())
        assert d < 1e-5, (name, d)
        report.append(f"mha[{name}]: q{tuple(q.shape)} k{tuple(k.shape)} "
                      f"mask{tuple(mask.shape) if mask is not None else None}: oracle vs reference max|diff| = {d:.1e}")
        gold[f"{name}_out_sub"] = y[..., ::8].numpy().copy()   # every 8th column keeps the fixture small
    np.savez_compressed(GOLD / "mha_cases.npz", names=np.array([c[0] for c in mha_cases()]), **gold)

    # ---- calculate_tiou + the averaging of inference.py:45-55
    cases = tiou_cases()
    per, mism = [], 0
    for gt, pred in cases:
        r = ref_tiou(gt, pred, THRESHOLDS)
        o = mmct.calculate_tiou(gt, pred, THRESHOLDS)
        mism += any(r[t] != o[t] for t in THRESHOLDS)
        per.append([r[t] for t in THRESHOLDS])
    assert mism == 0, f"{mism} tIoU cases differ from the reference"
    tIoU = {t: sum(p[i] for p in per) / len(per) for i, t in enumerate(THRESHOLDS)}   # inference.py:50-52
    average = sum(tIoU.values()) / len(tIoU)
    o_avg, o_by_t = mmct.atiou([c[0] for c in cases], [c[1] for c in cases], THRESHOLDS)
    assert o_avg == average and all(o_by_t[t] == tIoU[t] for t in THRESHOLDS)
    report.append(f"tiou: {len(cases)} cases (empty gt / empty pred / zero union / exact 1 included): oracle "
                  f"calculate_tiou identical to the reference; AtIoU {average:.12f} identical")
    np.savez_compressed(GOLD / "tiou_cases.npz", per_video=np.array(per, dtype=np.float64),
                        by_threshold=np.array([tIoU[t] for t in THRESHOLDS]), average=np.array(average),
                        n_cases=len(cases), seed=41)

    # ---- host loader vs RepurposeClipTest + collate_fn_test
    with tempfile.TemporaryDirectory() as td:
        ds_cfg = write_loader_dataset(Path(td))
        ref_ds = RepurposeClipTest(**ds_cfg)
        ours = read_test_set(ds_cfg)
        assert [e["video_id"] for e in ours] == [k["youtube_id"] for k in ref_ds.label], "availability filter differs"
        lg = {"names": np.array([e["video_id"] for e in ours])}
        for i, e in enumerate(ours):
            b = collate_fn_test([ref_ds[i]])                    # what inference.py's DataLoader yields
            v = load_video_features(*e["paths"], time_range=e["time_range"], n_labels=e["n_labels"])
            n = v["duration"]
            assert n == b["duration"][0] == b["visual_feats"].shape[1], (e["video_id"], n, b["duration"])
            assert np.array_equal(v["visual_feats"], b["visual_feats"][0].numpy())
            assert np.array_equal(v["audio_feats"], b["audio_feats"][0].numpy())
            txt = np.zeros((n, v["text_feats"].shape[1]), np.float32)   # the reference pads a short text track
            txt[:v["text_feats"].shape[0]] = v["text_feats"]
            assert np.array_equal(txt, b["text_feats"][0].numpy())
            assert e["gt_segments"] == b["gt_segments"][0]
            lg[f"{e['video_id']}_duration"] = n
            lg[f"{e['video_id']}_text_rows"] = v["text_feats"].shape[0]
            lg[f"{e['video_id']}_checksums"] = np.array([checksum(b["visual_feats"][0].numpy()),
                                                         checksum(b["audio_feats"][0].numpy()),
                                                         checksum(b["text_feats"][0].numpy())])
        report.append(f"loader: {len(ours)} videos (missing-modality video filtered like the reference); duration, "
                      "visual / audio / zero-padded text rows and gt_segments identical to RepurposeClipTest + "
                      f"collate_fn_test; durations {[int(lg[n + '_duration']) for n in lg['names']]}")
        np.savez_compressed(GOLD / "loader_cases.npz", **lg)

    pin = GOLD / "PIN_REPORT.txt"
    old = [l for l in pin.read_text().splitlines() if not l.startswith(("mha[", "tiou:", "loader:"))]
    pin.write_text("\n".join(old + report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
