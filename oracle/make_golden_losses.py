#!/usr/bin/env python
"""Pin oracle/losses.py against the reference and write tests/golden/losses_cases.npz.

Runs in the build container only (imports the unmodified reference from /root/reference):
`MMCTransformer.losses` (models/MMCTransformer.py:159-179) is evaluated on seeded inputs — random logits
over a wide range, binary labels, left-aligned validity masks, plus extreme logits — the oracle must agree
within fp32 summation-order noise, and the reference's value is stored as the golden number.
    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_losses.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"


def main(ref_dir="/root/reference"):
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref_dir)
    from models.MMCTransformer import MMCTransformer as RefModel   # the reference
    sys.path.insert(0, str(ROOT))
    from oracle import losses as oracle_losses

    cases, report = {}, []
    g = torch.Generator().manual_seed(123)
    specs = [("small", 2, 37, 4.0), ("batch", 5, 700, 6.0), ("extreme", 3, 129, 60.0), ("long", 2, 1801, 3.0)]
    for name, B, T, scale in specs:
        logits = torch.randn(B, T, 1, generator=g) * scale
        labels = (torch.rand(B, T, generator=g) < 0.2).float()
        lens = torch.randint(1, T + 1, (B,), generator=g)
        lens[0] = T
        masks = (torch.arange(T)[None, :] < lens[:, None])[:, None, :]
        logits = logits.requires_grad_(True)
        ref = RefModel.losses(None, masks, logits, None, labels, None, None)["cls_loss"]
        (ref / B).backward()                      # main.py:326-333: final_loss = cls_loss / batch_size
        cases[f"{name}_dlogits"] = logits.grad.numpy().copy()
        logits = logits.detach()
        ours = oracle_losses.losses(masks, logits, labels)
        dl = oracle_losses.losses_grad(masks, logits, labels, B)
        gerr = float((dl - torch.from_numpy(cases[f"{name}_dlogits"])).abs().max())
        assert gerr < 1e-6, (name, gerr)
        rel = abs(float(ours) - float(ref)) / max(abs(float(ref)), 1e-12)
        assert rel < 1e-6, (name, float(ref), float(ours))
        report.append(f"losses[{name}]: B={B} T={T} cls_loss={float(ref):.6f}; oracle rel. diff {rel:.1e}; "
                      f"d(cls_loss/B)/dlogits from the reference's autograd stored, oracle max|diff| {gerr:.1e}")
        cases[f"{name}_logits"] = logits.numpy()
        cases[f"{name}_labels"] = labels.numpy()
        cases[f"{name}_masks"] = masks.numpy()
        cases[f"{name}_loss"] = np.array(float(ref), dtype=np.float64)
    cases["names"] = np.array([s[0] for s in specs])
    np.savez_compressed(GOLD / "losses_cases.npz", **cases)
    pin = GOLD / "PIN_REPORT.txt"
    old = [l for l in pin.read_text().splitlines() if not l.startswith("losses[")]
    pin.write_text("\n".join(old + report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
