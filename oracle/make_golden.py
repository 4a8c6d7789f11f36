"""ORACLE (test infrastructure): pin the oracle against the REFERENCE ITSELF and write golden vectors.

Run in the build container, where the unmodified reference is mounted read-only:

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden [--ref /root/reference]

It imports the reference's own `models.MMCTransformer.MMCTransformer` and
`models.softnms.soft_nms_intervals_cpu`, runs them on seeded synthetic inputs (oracle/synth.py),
asserts that the restatements in oracle/ agree with them, and stores the reference's outputs under
tests/golden/ (small .npz files, committed).  The GPU box has no /root/reference; there the tests
compare the CUDA path with these fixtures and with the (now pinned) oracle.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    sys.dont_write_bytecode = True
    sys.path.insert(0, args.ref)
    from models.MMCTransformer import MMCTransformer as RefModel          # the reference
    from models.softnms import soft_nms_intervals_cpu as ref_soft_nms     # the reference
    sys.path.insert(0, str(ROOT))
    from oracle import mmct, synth
    from oracle.softnms import soft_nms_intervals_oracle
    from oracle.softnms_c import soft_nms_intervals_c
    from repurpose_b200.models.MMCTransformer import MMCTransformer as OurModel

    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    report = []

    # ------------------------------------------------------------------ forward + inference_
    torch.manual_seed(0)
    ref = RefModel(**synth.MODEL_CFG).eval()
    torch.manual_seed(0)
    ours = OurModel(**synth.MODEL_CFG)
    sd_ref = {k: v.clone() for k, v in ref.state_dict().items()}
    sd_ours = ours.state_dict()
    assert list(sd_ref.keys()) == list(sd_ours.keys()), "state-dict schema differs from the reference"
    for k in sd_ref:
        assert sd_ref[k].shape == sd_ours[k].shape and torch.equal(sd_ref[k], sd_ours[k]), k
    report.append(f"state dict: {len(sd_ref)} tensors identical to the reference under manual_seed(0)")

    lens = [700, 433]
    batch = synth.make_batch(lens, seed=11)
    fwd = {}
    for tag, sd in (("init", sd_ref), ("regbias", synth.bias_reg_head(sd_ref))):
        ref.load_state_dict(sd)
        with torch.no_grad():
            _, logits, offsets, _, _, feats = ref(batch)
        o_logits, o_offsets, o_feats = mmct.forward(sd, batch)
        def rel(a, b):
            return ((a - b).abs().max() / a.abs().max().clamp_min(1e-6)).item()
        d = max(rel(logits, o_logits), rel(offsets, o_offsets), rel(feats, o_feats))
        assert d < 1e-5, f"oracle forward deviates from the reference by {d} (relative to max|ref|)"
        report.append(f"forward[{tag}]: oracle vs reference max|diff|/max|ref| = {d:.2e} "
                      f"(logits {rel(logits, o_logits):.1e}, offsets {rel(offsets, o_offsets):.1e}, "
                      f"feats {rel(feats, o_feats):.1e})")
        fwd[f"{tag}_logits"] = logits.numpy()
        fwd[f"{tag}_offsets"] = offsets.numpy()
        fwd[f"{tag}_feats_sub"] = feats[:, :, ::16].numpy().copy()
        if tag == "regbias":
            # inference_ with the reference's CUDA semantics (clone-fed Soft-NMS, App. B.1)
            for i, vlen in enumerate(lens):
                r = ref.inference_single_video(batch["masks"][i], logits[i, :, 0], offsets[i],
                                               synth.TEST_CFG)
                ms = synth.max_seg_num(vlen, synth.TEST_CFG["max_seg_per_min"])
                keep = ref_soft_nms(r["scores"].clone(), r["segments"].clone(),
                                    sigma=synth.TEST_CFG["nms_sigma"],
                                    thresh=synth.TEST_CFG["min_score"], max_seg_num=ms)
                fwd[f"inf{i}_cand_segments"] = r["segments"].numpy()
                fwd[f"inf{i}_cand_scores"] = r["scores"].numpy()
                fwd[f"inf{i}_cand_labels"] = r["labels"].numpy()
                fwd[f"inf{i}_keep"] = np.asarray(keep, dtype=np.int64)
                fwd[f"inf{i}_segments"] = r["segments"][keep].numpy()
                fwd[f"inf{i}_scores"] = r["scores"][keep].numpy()
                fwd[f"inf{i}_labels"] = r["labels"][keep].numpy()
            o_inf = mmct.inference(sd, batch, synth.TEST_CFG)
            for i in range(len(lens)):
                assert np.array_equal(o_inf[i]["labels"].numpy(), fwd[f"inf{i}_labels"]), "oracle inference_"
                assert np.allclose(o_inf[i]["segments"].numpy(), fwd[f"inf{i}_segments"], atol=1e-4)
            report.append("inference_[regbias]: oracle kept labels identical to the reference "
                          f"({[len(fwd[f'inf{i}_keep']) for i in range(len(lens))]} segments, "
                          f"{[len(fwd[f'inf{i}_cand_scores']) for i in range(len(lens))]} candidates)")
    np.savez_compressed(GOLD / "forward_T700.npz", lens=np.array(lens), batch_seed=11, weight_seed=0,
                        **fwd)

    # ------------------------------------------------------------------ decode
    dec = {}
    g = torch.Generator().manual_seed(5)
    for ci, (T, vlen) in enumerate([(1801, 1801), (1801, 1200), (300, 300)]):
        logits = torch.randn(T, generator=g) * 2.0
        offsets = torch.rand(T, 2, generator=g) * 60.0
        mask = (torch.arange(T) < vlen)[None]
        r = ref.inference_single_video(mask, logits, offsets, synth.TEST_CFG)
        o = mmct.decode_single_video(mask, logits, offsets, synth.TEST_CFG)
        assert torch.equal(r["labels"], o["labels"]) and torch.equal(r["scores"], o["scores"])
        dec[f"c{ci}_logits"] = logits.numpy()
        dec[f"c{ci}_offsets"] = offsets.numpy()
        dec[f"c{ci}_len"] = vlen
        dec[f"c{ci}_segments"] = r["segments"].numpy()
        dec[f"c{ci}_scores"] = r["scores"].numpy()
        dec[f"c{ci}_labels"] = r["labels"].numpy()
        report.append(f"decode case {ci}: T={T} len={vlen} -> {r['scores'].numel()} candidates, oracle identical")
    dec["n_cases"] = 3
    np.savez_compressed(GOLD / "decode_cases.npz", **dec)

    # ------------------------------------------------------------------ Soft-NMS
    cases = []

    def add(name, scores, segs, sigma, thresh, max_seg):
        cases.append((name, np.asarray(scores, np.float32), np.asarray(segs, np.float32).reshape(-1, 2),
                      float(sigma), float(thresh), int(max_seg)))

    s, g_ = synth.make_candidates(1000, 1801, 1); add("n1000_k9", s, g_, 0.5, 0.01, 9)
    s, g_ = synth.make_candidates(4096, 8192, 2); add("n4096_k41", s, g_, 0.5, 0.01, 41)
    s, g_ = synth.make_candidates(50, 300, 3); add("n50_k20", s, g_, 0.5, 0.001, 20)
    s, g_ = synth.make_candidates(300, 600, 4); add("n300_noearlystop", s, g_, 0.5, 0.01, 300)
    s, g_ = synth.make_candidates(1, 100, 5); add("n1", s, g_, 0.5, 0.01, 5)
    s, g_ = synth.make_candidates(2, 100, 6); add("n2", s, g_, 0.5, 0.01, 5)
    add("n0", np.zeros(0), np.zeros((0, 2)), 0.5, 0.01, 5)
    s, g_ = synth.make_candidates(40, 200, 7); add("maxseg0", s, g_, 0.5, 0.01, 0)
    s, g_ = synth.make_candidates(64, 400, 8); add("ties", np.full(64, 0.75, np.float32), g_, 0.5, 0.01, 12)
    # heavy overlap: clusters of near-identical segments, so decayed scores fall below thresh and
    # the pre-swap-tscore / stale-length quirks decide the outcome
    rng = np.random.default_rng(9)
    centres = rng.choice([100.0, 400.0, 900.0, 1500.0], size=400)
    jitter = rng.normal(0, 1.5, size=(400, 2))
    segs = np.stack([centres - 25 + jitter[:, 0], centres + 25 + jitter[:, 1]], 1)
    sc = np.sort(rng.uniform(0.5, 1.0, 400).astype(np.float32))[::-1]
    add("clustered_thr0.01", sc, segs, 0.5, 0.01, 9)
    add("clustered_thr0.3", sc, segs, 0.5, 0.3, 30)
    add("clustered_sigma0.1", sc, segs, 0.1, 0.05, 60)
    s, g_ = synth.make_candidates(777, 1801, 10); add("unsorted", s[::-1].copy(), g_, 0.5, 0.01, 9)
    s, g_ = synth.make_candidates(1000, 1801, 12); add("n1000_k40", s, g_, 0.5, 0.01, 40)

    nms = {"names": np.array([c[0] for c in cases])}
    exp_mismatch = 0
    for name, s, g_, sigma, thresh, ms in cases:
        keep_ref = np.asarray(ref_soft_nms(torch.from_numpy(s.copy()), torch.from_numpy(g_.copy()),
                                           sigma=sigma, thresh=thresh, max_seg_num=ms), dtype=np.int64)
        keep_o, dsc_o = soft_nms_intervals_oracle(s, g_, sigma, thresh, ms, return_scores=True)
        assert np.array_equal(keep_ref, keep_o), f"oracle Soft-NMS differs from the reference on {name}"
        keep_c, dsc_c = soft_nms_intervals_c(s, g_, sigma, thresh, ms, return_scores=True)
        same_c = np.array_equal(keep_ref, keep_c)
        exp_mismatch += 0 if same_c else 1
        dmax = float(np.abs(dsc_o - dsc_c).max()) if same_c and len(dsc_o) else 0.0
        report.append(f"softnms[{name}]: N={len(s)} M={ms} -> kept {len(keep_ref)}; numpy-oracle == ref; "
                      f"C-oracle {'==' if same_c else '!='} ref (max|dscore diff| {dmax:.1e})")
        nms[f"{name}_scores"] = s
        nms[f"{name}_segments"] = g_
        nms[f"{name}_params"] = np.array([sigma, thresh, ms], dtype=np.float64)
        nms[f"{name}_keep"] = keep_ref
        nms[f"{name}_dscores"] = dsc_o
    np.savez_compressed(GOLD / "softnms_cases.npz", **nms)
    report.append(f"C oracle keep mismatches vs reference: {exp_mismatch} of {len(cases)} cases")

    (GOLD / "PIN_REPORT.txt").write_text("\n".join(report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
