"""ORACLE (test infrastructure): stages the UNMODIFIED reference implementation of the hot path into
oracle/_ref/ so that it can be timed on the GPU box's host cores (`bench.py --impl reference`,
`cpu_baseline.kind = "reference"`).

    python -m oracle.build_ref            # also run by __graft_entry__.build() and `make -C oracle ref`

The reference is pure Python, so "building" it is a byte-for-byte staging of the five modules the path
lives in (models/MMCTransformer.py, models/softnms.py, models/losses.py, models/transformer.py,
utils/metrics.py) plus configs/Repurpose.yaml, read where they lie under /root/reference (or $REPURPOSE_REF).
oracle/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored, so the staged
copy travels to the GPU box like the built .so files.  Nothing in the product path imports it; a MANIFEST
with sha256 digests records exactly what was staged.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys
from pathlib import Path

_DIR = Path(__file__).resolve().parent
REF_OUT = _DIR / "_ref"
FILES = ["models/__init__.py", "models/MMCTransformer.py", "models/softnms.py", "models/losses.py",
         "models/transformer.py", "utils/metrics.py", "configs/Repurpose.yaml"]


def reference_root() -> Path | None:
    for cand in (os.environ.get("REPURPOSE_REF"), "/root/reference"):
        if cand and (Path(cand) / "models" / "MMCTransformer.py").exists():
            return Path(cand)
    return None


def build() -> Path | None:
    """Stages the reference files; returns oracle/_ref (or None when no reference tree is present, e.g. on the
    GPU box, where the copy staged in the build container is used as is)."""
    root = reference_root()
    if root is None:
        return REF_OUT if available() else None
    lines = []
    for rel in FILES:
        src, dst = root / rel, REF_OUT / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        if src.exists():
            shutil.copyfile(src, dst)
            lines.append(f"{hashlib.sha256(dst.read_bytes()).hexdigest()}  {rel}")
        elif rel.endswith("__init__.py"):
            dst.write_text("")
    (REF_OUT / "utils" / "__init__.py").touch()
    (REF_OUT / "MANIFEST.txt").write_text(f"staged unmodified from {root}\n" + "\n".join(lines) + "\n")
    return REF_OUT


def available() -> bool:
    return (REF_OUT / "models" / "MMCTransformer.py").exists() and (REF_OUT / "models" / "softnms.py").exists()


def import_reference():
    """Returns the reference's `models.MMCTransformer` module imported from oracle/_ref (None if not staged)."""
    if not available():
        return None
    import importlib
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
    sys.path.insert(0, str(REF_OUT))
    try:
        mod = importlib.import_module("models.MMCTransformer")
    finally:
        sys.path.remove(str(REF_OUT))
        # keep the reference's package under a private name so it cannot shadow anything else
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            sys.modules["_rp_ref_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)
    return mod


if __name__ == "__main__":
    print(build())
