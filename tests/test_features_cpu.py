"""CPU suite, part 4: feature-file loading and ragged collation (host side of the GPU collate)."""
import numpy as np
import torch

from repurpose_b200.features import collate_ragged, load_video_features


def test_load_video_features_follows_reference_slicing(tmp_path):
    rng = np.random.default_rng(0)
    vis = rng.normal(size=(50, 8)).astype(np.float64)     # extractors may write float64
    aud = rng.normal(size=(48, 16)).astype(np.float32)    # audio one step shorter than visual
    txt = rng.normal(size=(45, 4)).astype(np.float32)     # text shorter still: NOT part of the minimum
    for name, a in (("v", vis), ("a", aud), ("t", txt)):
        np.save(tmp_path / f"{name}.npy", a)
    paths = [tmp_path / "v.npy", tmp_path / "a.npy", tmp_path / "t.npy"]
    full = load_video_features(*paths, time_range=[0, 50.0])
    assert full["duration"] == 48 and full["visual_feats"].dtype == np.float32
    assert full["text_feats"].shape[0] == 45                 # dataset/RepurposeClip.py:975-980
    np.testing.assert_array_equal(full["visual_feats"], vis[:48].astype(np.float32))
    part = load_video_features(*paths, time_range=[10.7, 30.2], n_labels=15)
    assert part["duration"] == 15                            # rows [10, 30) then min with the label count
    np.testing.assert_array_equal(part["audio_feats"], aud[10:25])


def test_collate_ragged_layout():
    vids = [{"visual_feats": torch.full((3, 8), 1.0), "audio_feats": torch.full((3, 16), 2.0),
             "text_feats": torch.full((2, 8), 3.0)},
            {"visual_feats": torch.full((5, 8), 4.0), "audio_feats": torch.full((5, 16), 5.0),
             "text_feats": torch.full((5, 8), 6.0), "video_id": "b"}]
    b = collate_ragged(vids, pin=False)
    assert b["ragged"] and b["duration"] == [3, 5] and b["video_id"] == [0, "b"]
    assert b["visual_feats"].shape == (8, 8) and b["text_feats"].shape == (7, 8)
    assert b["row_offsets"].tolist() == [0, 3] and b["text_offsets"].tolist() == [0, 2]
    assert b["text_lens"].tolist() == [2, 5] and b["lens"].tolist() == [3, 5]
    assert b["visual_feats"][3:].eq(4.0).all() and b["text_feats"][:2].eq(3.0).all()


def test_affinity_helpers_are_best_effort():
    from repurpose_b200.affinity import _parse_cpulist, bind_to_gpu_numa
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11} and _parse_cpulist("") == set()
    info = bind_to_gpu_numa(0)              # no GPU / no sysfs here: must not raise, must not change anything
    assert set(info) == {"numa_node", "cpus"}


def test_load_video_features_bf16(tmp_path):
    rng = np.random.default_rng(1)
    arrs = {"v": rng.normal(size=(20, 8)).astype(np.float32), "a": rng.normal(size=(20, 16)).astype(np.float32),
            "t": rng.normal(size=(18, 8)).astype(np.float32)}
    for k, a in arrs.items():
        np.save(tmp_path / f"{k}.npy", a)
    paths = [tmp_path / "v.npy", tmp_path / "a.npy", tmp_path / "t.npy"]
    v = load_video_features(*paths, dtype="bf16")
    assert v["visual_feats"].dtype == torch.bfloat16 and v["duration"] == 20 and v["text_feats"].shape == (18, 8)
    assert torch.equal(v["audio_feats"], torch.from_numpy(arrs["a"]).to(torch.bfloat16))   # round to nearest even
    from repurpose_b200.features import ragged_batch
    b = ragged_batch([v, v])
    assert b["parts"]["visual_feats"][0].dtype == torch.bfloat16 and b["text_lens"].tolist() == [18, 18]
    import pytest
    with pytest.raises(ValueError):
        load_video_features(*paths, dtype="fp16")


def test_infer_read_test_set_follows_reference_rules(tmp_path):
    import json
    from repurpose_b200.infer import read_test_set
    dirs = {k: tmp_path / k for k in ("video_path", "audio_path", "text_path")}
    for d in dirs.values():
        d.mkdir()
    for vid in ("a", "b"):                       # "c" lacks its audio file -> filtered out like the reference does
        for d in dirs.values():
            np.save(d / f"{vid}.npy", np.zeros((5, 4), dtype=np.float32))
    np.save(dirs["video_path"] / "c.npy", np.zeros((5, 4), dtype=np.float32))
    labels = [{"youtube_id": "a", "timeRange": [0, 99.5], "timeRangeOffset": [0, 99.5], "segmentsOffset": [[1.0, 20.0]]},
              {"youtube_id": "c", "timeRange": [0, 10], "timeRangeOffset": [0, 10], "segmentsOffset": []},
              {"youtube_id": "b", "timeRange": [30, 80.2], "timeRangeOffset": [0, 50.2], "segmentsOffset": [[2.0, 30.0], [31.0, 45.0]]}]
    (tmp_path / "test.json").write_text(json.dumps(labels))
    cfg = {"label_path": str(tmp_path / "test.json"), **{k: str(v) for k, v in dirs.items()}}
    es = read_test_set(cfg)
    assert [e["video_id"] for e in es] == ["a", "b"]
    assert es[0]["n_labels"] == 100 and es[1]["n_labels"] == 51          # int(t1 - t0) + 1
    assert es[1]["time_range"] == [30, 80.2] and es[1]["gt_segments"] == [[2.0, 30.0], [31.0, 45.0]]
    assert es[0]["paths"][1].endswith("audio_path/a.npy")


def test_loader_matches_reference_dataset_golden(tmp_path, golden_dir):
    """read_test_set + load_video_features vs the reference's RepurposeClipTest + collate_fn_test
    (dataset/RepurposeClip.py:578-606, 962-1038) on the dataset oracle/make_golden_aux.py pinned:
    same availability filter, durations, feature rows and zero-padded text rows."""
    import numpy as np
    from oracle.make_golden_aux import checksum, write_loader_dataset
    from repurpose_b200.features import load_video_features
    from repurpose_b200.infer import read_test_set
    g = np.load(golden_dir / "loader_cases.npz")
    entries = read_test_set(write_loader_dataset(tmp_path))
    assert [e["video_id"] for e in entries] == g["names"].tolist()
    for e in entries:
        v = load_video_features(*e["paths"], time_range=e["time_range"], n_labels=e["n_labels"])
        name, n = e["video_id"], v["duration"]
        assert n == int(g[f"{name}_duration"]) and v["text_feats"].shape[0] == int(g[f"{name}_text_rows"])
        txt = np.zeros((n, v["text_feats"].shape[1]), np.float32)
        txt[:v["text_feats"].shape[0]] = v["text_feats"]
        got = np.array([checksum(v["visual_feats"]), checksum(v["audio_feats"]), checksum(txt)])
        assert np.allclose(got, g[f"{name}_checksums"], rtol=0, atol=1e-6), name
