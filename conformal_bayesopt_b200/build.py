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


TORCH_SRC = os.path.join(HERE, "csrc", "torch_ext.cpp")
TORCH_LIB = os.path.join(LIB_DIR, "libcbo_torch.so")


def sources():
    d = os.path.join(HERE, "csrc")
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f != "torch_ext.cpp"] + \
        [os.path.join(os.path.dirname(HERE), "include", "cbo.h")]


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


def torch_ext_is_stale() -> bool:
    if not os.path.exists(TORCH_LIB):
        return True
    t = os.path.getmtime(TORCH_LIB)
    return any(os.path.getmtime(f) > t for f in (TORCH_SRC, os.path.join(os.path.dirname(HERE), "include", "cbo.h")))


def build_torch_ext(force: bool = False) -> str:
    """Compile csrc/torch_ext.cpp (TORCH_LIBRARY(cbo, ...): a shim over the C ABI, no kernels) into lib/libcbo_torch.so,
    linked against lib/libcbo.so (rpath $ORIGIN) and the torch libraries of this interpreter."""
    if not force and not torch_ext_is_stale():
        return TORCH_LIB
    build()
    import torch
    from torch.utils import cpp_extension as ce
    cxx = shutil.which("g++")
    if cxx is None:
        raise NvccNotFound("g++ not found: cannot build libcbo_torch.so")
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    for inc in ce.include_paths() + ["/usr/local/cuda/include"]:
        cmd += ["-isystem", inc]
    cmd += ["-o", TORCH_LIB, TORCH_SRC, "-L", LIB_DIR, "-lcbo", "-L", tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda",
            "-ltorch", "-L", "/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + tlib]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    return TORCH_LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_torch_ext(force=True))
