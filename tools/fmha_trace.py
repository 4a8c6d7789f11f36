"""Debug: in-kernel clock trace of one attention CTA (tools/build_variant.sh trace fmha.cu -DRP_FMHA_TRACE, RP_LIB_PATH=ab/lib_trace.so)."""
import ctypes as C, os, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from repurpose_b200 import _lib
from repurpose_b200._lib import check, cur_stream, ptr
lib = _lib.load()
B, T, H, D = int(os.environ.get("TRACE_B", 32)), 1801, 8, 512
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
f = lambda r, i: int(tr[r, i] - t0) if tr[r, i] > 0 else -1
print("j | PV.h0 PV.h1 issue | QK(j+2).h0 .h1 issue | softmax: loop top, steps done | (iteration length)")
for j in range(14):
    print(j, "|", f(0, j), f(1, j), "|", f(2, j), f(3, j), "|", f(4, 2 * j), f(4, 2 * j + 1), "|",
          f(4, 2 * j + 2) - f(4, 2 * j))
print("per-step start times relative to loop top (steps 0..8), iterations 4..7")
for j in range(4, 8):
    print(j, [f(5, 16 * j + s) - f(4, 2 * j) for s in range(9)], "tail", f(4, 2 * j + 2) - f(4, 2 * j + 1))
print("step 3 detail: [before ld_wait, after, after max | D: start, before STTM, after STTM]; tail: [start, after st_wait, after ld_wait, after max, after arrives]  (relative to loop top)")
for j in range(4, 8):
    b = f(4, 2 * j)
    print(j, [f(6, 16 * j + k) - b for k in (0, 1, 2, 8, 9, 10)], [f(6, 16 * j + k) - b for k in (3, 4, 5, 6, 7)], "next top", f(4, 2 * j + 2) - b)
print("j | softmax top | PV(j).h0 PV(j).h1 issue | QK(j+1).h0 .h1 issue (needed at step 1 / 5 of iteration j)")
for j in range(3, 10):
    print(j, "|", f(4, 2 * j), "|", f(0, j), f(1, j), "|", f(2, j + 1), f(3, j + 1))
