"""repurpose_b200 — B200 (sm_100a) implementation of the Repurpose inference hot path
(MMCTransformer forward -> per-video decode -> Gaussian Soft-NMS) behind the reference's Python
call signatures.  See DESIGN.md for the path/boundary and include/repurpose_b200.h for the C ABI."""
__version__ = "0.1.0"

from . import _lib  # noqa: F401


def __getattr__(name):  # lazy: keep `import repurpose_b200` cheap for the CPU-only checks
    if name == "MMCTransformer":
        from .models.MMCTransformer import MMCTransformer
        return MMCTransformer
    if name in ("soft_nms_intervals_cpu", "soft_nms_intervals", "soft_nms_batched"):
        from .models import softnms
        return getattr(softnms, name)
    if name == "MultiHeadAttention":
        from .models.transformer import MultiHeadAttention
        return MultiHeadAttention
    raise AttributeError(name)
