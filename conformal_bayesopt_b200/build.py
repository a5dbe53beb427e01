"""In-tree build of libcbo.so (nvcc, sm_100a).  The .so is git-ignored but travels with gpurun."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "cbo_api.cu")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libcbo.so")
# -cudart shared: the library resolves the CUDA runtime from the process (torch has loaded libcudart.so.12 already; the
# rpath covers a bare ctypes load) instead of embedding a static copy of it.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "--shared", "-Xcompiler", "-fPIC", "-cudart", "shared", "-Xlinker", "-rpath,/usr/local/cuda/lib64",
]


class NvccNotFound(RuntimeError):
    """No CUDA compiler on this host (the only build failure `_lib.load()` tolerates, and only if a built .so exists)."""


def sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d))] + [os.path.join(os.path.dirname(HERE), "include", "cbo.h")]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/cbo_api.cu (which includes the .cuh kernels) into lib/libcbo.so."""
    if not force and not is_stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise NvccNotFound("nvcc not found: cannot build libcbo.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
