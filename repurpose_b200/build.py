"""In-tree build of librepurpose_b200.so (hand-written sm_100a kernels + the C ABI).

    python -m repurpose_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU; the resulting .so sits next to this file so that it
travels with the source tree (it is git-ignored).  cudart is linked statically and the driver API
is resolved at run time, so the library also dlopen()s on a CPU-only box (symbol checks).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "_build"
LIB_PATH = PKG_DIR / "librepurpose_b200.so"
SOURCES = ["host_util.cu", "gemm.cu", "fmha.cu", "fmha_bwd.cu", "rowwise.cu", "decode_nms.cu", "metrics.cu", "train.cu", "api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


NVCC_FLAGS += os.environ.get("RP_EXTRA_NVCC_FLAGS", "").split()


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if Path(cand).exists() else "nvcc"


def _source_digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h"))
                    + [PKG_DIR.parent / "include" / "repurpose_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    BUILD_DIR.mkdir(exist_ok=True)
    stamp = BUILD_DIR / "digest.txt"
    digest = _source_digest()
    if not force and LIB_PATH.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB_PATH
    nvcc = _nvcc()

    def compile_one(src: str):
        obj = BUILD_DIR / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    log = []
    for src, obj, r in results:
        log.append(f"==== {src}\n{r.stdout}{r.stderr}")
        if r.returncode != 0:
            sys.stderr.write(log[-1])
            raise RuntimeError(f"nvcc failed on {src}")
    (BUILD_DIR / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    objs = [str(obj) for _, obj, _ in results]
    cmd = [nvcc, "-shared", "-o", str(LIB_PATH), *objs, "-cudart", "static",
           "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    stamp.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
