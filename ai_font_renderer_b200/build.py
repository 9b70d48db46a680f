"""Builds libafr_sm100.so (the C-ABI CUDA library of the hot path) in-tree with nvcc for sm_100a.

`python -m ai_font_renderer_b200.build` or `__graft_entry__.build()`. nvcc cross-compiles
without a GPU; the .so is git-ignored but travels to the GPU box with the working tree.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libafr_sm100.so")
STAMP_PATH = LIB_PATH + ".stamp"
SOURCES = ["afr_api.cu", "afr_gemm.cu", "afr_frontend.cu", "afr_wide.cu", "afr_elementwise.cu"]
HEADERS = ["afr_ptx.cuh", "afr_gemm.cuh", "afr_philox.cuh", "afr_internal.h", "../../include/afr_sm100.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
] + os.environ.get("AFR_EXTRA_NVCC_FLAGS", "").split()   # e.g. -DAFR_PHASE_TIMING (tools/phase_timing.py)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libafr_sm100.so")


def source_digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as f:
        return f.read().strip() == source_digest()


def build(force: bool = False, verbose: bool = True) -> str:
    """Compile the library if sources changed. Returns the path of the .so.

    Safe under `torchrun` (every rank may call it at once): the build runs under an exclusive file
    lock, nvcc writes to a temporary file that is renamed over the library only when complete, and
    a rank that waited for the lock re-checks the stamp instead of compiling again -- nobody can
    dlopen a half-written library."""
    if not force and is_current():
        return LIB_PATH
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():          # another process built it while we waited
                return LIB_PATH
            tmp = f"{LIB_PATH}.tmp.{os.getpid()}"
            cmd = [_nvcc(), *NVCC_FLAGS, "-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            if verbose:
                print("[afr build]", " ".join(cmd).replace(tmp, LIB_PATH), flush=True)
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError("nvcc failed building libafr_sm100.so")
            os.replace(tmp, LIB_PATH)
            with open(STAMP_PATH + ".tmp", "w") as f:
                f.write(source_digest())
            os.replace(STAMP_PATH + ".tmp", STAMP_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
