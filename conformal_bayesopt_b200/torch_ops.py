"""`torch.ops.cbo.*`: the thin torch C++ extension over the C ABI (csrc/torch_ext.cpp, lib/libcbo_torch.so).

The operators are dispatcher-registered (TORCH_LIBRARY), `conformal_acq` with an Autograd kernel, so they compose with
torch.autograd / torch.no_grad like any ATen op.  `load()` builds the shim in-tree when its source is newer (g++ only: it
contains no kernels) and loads it; the Python classes use it for the plain cUCB / cEI call and keep the ctypes binding
(`_lib.py`) for everything that needs the raw ABI (extra outputs, refinement hooks, cNEI, host-buffer pipeline)."""
from __future__ import annotations

import os

import torch

from . import _lib
from . import build as _build

_OPS = None


def load():
    """torch.ops.cbo with the library loaded (raises if it cannot be built or loaded: never a silent fallback)."""
    global _OPS
    if _OPS is not None:
        return _OPS
    _lib.load()                                   # libcbo.so first: the shim links against it
    path = os.environ.get("CBO_TORCH_LIB_PATH") or _build.TORCH_LIB
    if path == _build.TORCH_LIB and _build.torch_ext_is_stale():
        try:
            _build.build_torch_ext()
        except _build.NvccNotFound as e:
            if not os.path.exists(path):
                raise _lib.CboError(f"libcbo_torch.so is missing and could not be built: {e}") from e
    torch.ops.load_library(path)
    _OPS = torch.ops.cbo
    return _OPS


KIND = {"ucb": 0, "ei": 1}


def state_args(state):
    """The model-state leading arguments of torch.ops.cbo.conformal_acq from a GPModelState."""
    return (state.X, state.alpha, state.linv, state.linv_t, state.inv_lengthscale, float(state.mean), float(state.outputscale),
            float(state.noise))


def conformal_acq(state, Xq, *, eps_grid, z, eps_mc, u_rand=None, ratio=None, kind="ucb", alpha, temp, grid_res,
                  max_grid_refinements=0, optimistic=False, randomized=False, beta_prime=0.0, best_f=0.0, chunk=0):
    """torch.ops.cbo.conformal_acq: acq (b,), differentiable in Xq (b, q, d) through the registered Autograd kernel."""
    ops = load()
    rw = [] if ratio is None else [ratio.W1, ratio.b1, ratio.W2, ratio.b2, ratio.W3, ratio.b3]
    try:
        return ops.conformal_acq(*state_args(state), Xq, eps_grid, z, eps_mc, u_rand, rw, KIND[kind], float(alpha), float(temp),
                                 int(grid_res), int(max_grid_refinements), bool(optimistic), bool(randomized), float(beta_prime),
                                 float(best_f), int(chunk))
    except NotImplementedError:
        raise
    except RuntimeError as e:                      # TORCH_CHECK failures of the shim: the same error type as the ctypes binding
        raise _lib.CboError(str(e).split("\n")[0]) from e
