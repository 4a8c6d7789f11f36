"""Debug: dump the in-kernel clock trace of one FMHA CTA (build with RP_EXTRA_NVCC_FLAGS=-DRP_FMHA_TRACE)."""
import ctypes as C, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from repurpose_b200 import _lib
from repurpose_b200._lib import check, cur_stream, ptr
lib = _lib.load()
import os
B, T, H, D = int(os.environ.get('TRACE_B', 4)), 1801, 8, 512
qkv = torch.randn(B, T, 3 * D, device="cuda"); qkv[..., :D] *= 1.4427 / 8; qkv = qkv.bfloat16()
o = torch.empty(B, T, D, dtype=torch.bfloat16, device="cuda")
for _ in range(2):
    check(lib.rp_fmha(ptr(qkv), ptr(qkv) + D * 2, ptr(qkv) + 4 * D, ptr(o), 3 * D, 3 * D, 3 * D, D, T * 3 * D,
                      T * 3 * D, T * 3 * D, T * D, B, H, T, T, 0, 0, 0, 0, 0, cur_stream()), "fmha")
torch.cuda.synchronize()
buf = np.zeros(8 * 512, dtype=np.uint64)
raw = C.CDLL(str(_lib.LIB_PATH))
raw.rp_debug_fmha_trace.argtypes = [C.c_void_p]
assert raw.rp_debug_fmha_trace(buf.ctypes.data) == 0
tr = buf.reshape(8, 512).astype(np.int64)
t0 = tr[tr > 0].min()
n = 8
print("j | QKissue(q0,q1 for j+1) | PVissue(q0,q1) | WG0: wait_s, got_s, ldtm, max, xchg, got_pv, A_issued, B_done | WG1: ...")
for j in range(n):
    f = lambda r, i: int(tr[r, i] - t0) if tr[r, i] > 0 else -1
    print(j, "|", f(0, j), f(1, j), "|", f(2, j), f(3, j), "|", [f(4, 8 * j + k) for k in range(8)], "|",
          [f(5, 8 * j + k) for k in range(8)])
