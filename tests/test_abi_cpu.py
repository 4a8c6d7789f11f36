"""CPU suite, part 2: the C-ABI library loads on a box without a GPU, exports every symbol that
include/repurpose_b200.h declares, and the product path fails loudly instead of falling back."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

from repurpose_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "repurpose_b200.h"


def _declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(rp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and header disagree"
    assert lib.rp_abi_version() == 1


def test_library_has_no_link_time_driver_dependency():
    # must dlopen on a CPU-only box: the driver API is resolved at run time
    import subprocess
    out = subprocess.run(["ldd", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out
    assert "not found" not in out


def test_product_path_never_imports_the_oracle():
    for p in (ROOT / "repurpose_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly():
    lib = _lib.load()
    cfg = _lib.RpModelCfg(512, 2048, 384, 512, 16, 8, 2048, 256, 5000)
    h = C.c_void_p()
    rc = lib.rp_create(C.byref(cfg), C.byref(h))
    assert rc != 0 and lib.rp_last_error()
    from repurpose_b200.models.MMCTransformer import MMCTransformer
    from repurpose_b200.models.softnms import soft_nms_intervals_cpu
    from oracle import synth
    torch.manual_seed(0)
    m = MMCTransformer(512, 2048, 384, 512, 1, 3, 3, 8)
    with pytest.raises(_lib.RepurposeError):
        m(synth.make_batch([32], seed=0))
    with pytest.raises(_lib.RepurposeError):
        soft_nms_intervals_cpu(torch.rand(4), torch.rand(4, 2))


def test_invalid_arguments_return_status_not_abort():
    lib = _lib.load()
    assert lib.rp_gemm_bf16(0, None, 0, None, 0, None, 0, None, None, 0, 1, 256, 64, None) != 0
    assert b"null" in lib.rp_last_error()
    assert lib.rp_workspace_bytes(None, 1, 1) == -1
    assert lib.rp_create(None, None) != 0


def test_module_is_copyable_and_picklable():
    """copy.deepcopy / torch.save(model) as EMA and checkpoint-whole-module code does: the native
    handle (a ctypes pointer) must stay behind and the copy repacks lazily."""
    import copy
    import ctypes
    import io
    import torch
    from repurpose_b200 import synth
    from repurpose_b200.models.MMCTransformer import MMCTransformer
    m = MMCTransformer(**dict(synth.MODEL_CFG, self_num_layers=1))
    m._handle = ctypes.c_void_p(1234)          # as after a first forward
    try:
        m2 = copy.deepcopy(m)
        buf = io.BytesIO()
        torch.save(m, buf)
        buf.seek(0)
        m3 = torch.load(buf, weights_only=False)
    finally:
        m._handle = None                        # nothing for __del__ to destroy
    for c in (m2, m3):
        assert c._handle is None and c._weights_sig is None and c._workspace is None
        assert all(torch.equal(a, b) for a, b in zip(c.state_dict().values(), m.state_dict().values()))
