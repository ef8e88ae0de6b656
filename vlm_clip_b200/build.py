"""Builds libvlmclip_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Run as `python -m vlm_clip_b200.build` or through `__graft_entry__.build()`.  The library links the static
CUDA runtime, so it has no link-time dependency on libcuda and can be dlopen'ed on a box without a GPU
(the non-GPU tests check that every symbol of include/vlmclip.h is exported).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "build"
LIB = LIBDIR / "libvlmclip_b200.so"

SOURCES = ["api.cu", "gemm_tcgen05.cu", "rowwise.cu", "attention.cu", "attention_tc.cu", "attention_pp.cu", "attention_1q.cu", "encoder.cu", "preprocess.cu", "adapter.cu", "heads.cu", "clip_loss.cu", "optim.cu", "backward.cu",
           "attention_bwd.cu", "attention_bwd_mma.cu", "shared_adapter.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA kernels cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "vlmclip.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(nvcc: str, src: str, log: list) -> Path:
    obj = OBJDIR / (Path(src).stem + ".o")
    cmd = [nvcc, *NVCC_FLAGS, "-I", str(ROOT / "include"), "-c", str(CSRC / src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append((src, r.stdout + r.stderr))
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    stamp = LIBDIR / ".build_digest"
    digest = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        return LIB
    nvcc = _nvcc()
    log: list = []
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s, log), SOURCES))
    link = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-cudart", "static",
            "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    (LIBDIR / "ptxas_info.txt").write_text("\n".join(f"==== {s}\n{t}" for s, t in log))
    stamp.write_text(digest)
    if verbose:
        for s, t in log:
            print(f"==== {s}\n{t}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
