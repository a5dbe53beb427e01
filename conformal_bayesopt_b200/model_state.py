"""Plain-tensor hand-off of an exact GP to the kernels (SURVEY.md §8b "Model-state hand-off").

The reference passes a BoTorch `SingleTaskGP` object into its acquisition classes
(conformalbo/utils.py:382-390, 440-458); the kernels need, per output, the training inputs, the
hyper-parameters, alpha = A^-1 (y - m) and the explicit inverse Cholesky factor L^-1 of
A = K + noise I (the same quantity GPyTorch's `root_inv_decomposition` holds for the reference's
fantasy update, App. B-4).  Built once per model fit by libcbo's own builder
(`cbo_model_build_f64`: kernel matrix, blocked Cholesky and triangular inverse on the fp64 tensor
path); no torch.linalg on the product path, no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional

import torch

from . import _lib

SQRT5 = math.sqrt(5.0)


def matern52(x1: torch.Tensor, x2: torch.Tensor, lengthscale: torch.Tensor, outputscale: float) -> torch.Tensor:
    """ScaleKernel(MaternKernel(nu=2.5, ARD)) -- BoTorch SingleTaskGP's default covariance (App. B-1)."""
    a = x1 / lengthscale
    b = x2 / lengthscale
    r = torch.cdist(a, b, compute_mode="donot_use_mm_for_euclid_dist").clamp_min(1e-15)
    return outputscale * (1.0 + SQRT5 * r + (5.0 / 3.0) * r * r) * torch.exp(-SQRT5 * r)


@dataclass
class GPModelState:
    """One output's exact GP as device tensors (fp64), padded to Np = multiple of 128."""
    X: torch.Tensor                # (N, d)
    y: torch.Tensor                # (N,)
    mean: float
    lengthscale: torch.Tensor      # (d,)
    outputscale: float
    noise: float
    alpha: torch.Tensor = field(repr=False, default=None)     # (Np,)
    linv: torch.Tensor = field(repr=False, default=None)      # (Np, Np) lower
    linv_t: torch.Tensor = field(repr=False, default=None)    # (Np, Np) upper
    inv_lengthscale: torch.Tensor = field(repr=False, default=None)

    @property
    def N(self) -> int:
        return self.X.shape[0]

    @property
    def d(self) -> int:
        return self.X.shape[1]

    @property
    def Np(self) -> int:
        return self.alpha.shape[0]

    @property
    def device(self):
        return self.X.device

    @classmethod
    def fit_fixed(cls, X, y, mean, lengthscale, outputscale, noise, device=None, check_pd: bool = True) -> "GPModelState":
        """Build the state for FIXED hyper-parameters (hyper-parameter fitting is out of scope)."""
        device = torch.device(device) if device is not None else X.device
        X = X.to(device=device, dtype=torch.float64).contiguous()
        y = y.to(device=device, dtype=torch.float64).reshape(-1).contiguous()
        N, d = X.shape
        ls = torch.as_tensor(lengthscale, dtype=torch.float64, device=device).reshape(-1)
        if ls.numel() == 1:
            ls = ls.expand(d)
        ls = ls.contiguous()
        if not X.is_cuda:
            raise _lib.CboError("libcbo has no CPU path: the model state is built on a CUDA device")
        lib = _lib.load()
        Np = (N + 127) // 128 * 128
        inv_ls = (1.0 / ls).contiguous()
        alpha = torch.empty(Np, dtype=torch.float64, device=device)
        linv = torch.empty(Np, Np, dtype=torch.float64, device=device)
        linv_t = torch.empty(Np, Np, dtype=torch.float64, device=device)
        info = torch.zeros(1, dtype=torch.int32, device=device)
        nbytes = lib.cbo_model_workspace_bytes(N, d)
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        ws_ptr = (ws.data_ptr() + 255) // 256 * 256
        st = lib.cbo_model_build_f64(X.data_ptr(), y.data_ptr(), N, d, inv_ls.data_ptr(), float(mean), float(outputscale),
                                     float(noise), alpha.data_ptr(), linv.data_ptr(), linv_t.data_ptr(), info.data_ptr(),
                                     ws_ptr, nbytes, torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "cbo_model_build_f64")
        if check_pd and int(info.item()) != 0:        # one host sync per model fit
            raise _lib.CboError(f"K + noise I is not positive definite (pivot of row {int(info.item()) - 1})")
        return cls(X=X, y=y, mean=float(mean), lengthscale=ls, outputscale=float(outputscale), noise=float(noise),
                   alpha=alpha, linv=linv, linv_t=linv_t, inv_lengthscale=inv_ls)

    @classmethod
    def from_botorch(cls, model, output_index: int = 0) -> "GPModelState":
        """Duck-typed extraction from a BoTorch 0.7 `SingleTaskGP`.  botorch is not needed
        to import this module; the attributes below are the ones the reference itself reads."""
        for attr in ("input_transform", "outcome_transform"):
            if getattr(model, attr, None) is not None:
                raise NotImplementedError(f"the model carries an {attr}: the kernels take the GP in its own (transformed) "
                                          "space only -- build the GPModelState from the transformed data explicitly")
        X = model.train_inputs[0]
        y = model.train_targets
        ls = model.covar_module.base_kernel.lengthscale
        os_ = model.covar_module.outputscale
        noise = model.likelihood.noise
        mean = model.mean_module.constant
        if X.ndim == 3:  # batched multi-output model: (t, N, d)
            X, y = X[output_index], y[output_index]
            ls, os_, noise, mean = ls[output_index], os_[output_index], noise[output_index], mean[output_index]
        def scalar(v):       # never through torch.as_tensor(python float): that rounds to float32
            return float(v.detach().reshape(-1)[0]) if torch.is_tensor(v) else float(v)
        return cls.fit_fixed(X.detach(), y.detach(), scalar(mean), ls.detach().reshape(-1), scalar(os_), scalar(noise),
                             device=X.device)

    def append(self, X_new: torch.Tensor, y_new: torch.Tensor, check_pd: bool = True) -> "GPModelState":
        """A NEW state with the observations (X_new (q, d), y_new (q)) added, hyper-parameters unchanged: rank-1 extensions
        of L^-1 and alpha by `cbo_model_append_f64`, O(q N^2) instead of the O(N^3) rebuild.  This state is left untouched
        (the factors are copied first: 2 x 8 MB at N = 1024, microseconds on HBM3e)."""
        lib = _lib.load()
        dev = self.device
        X_new = X_new.to(device=dev, dtype=torch.float64).reshape(-1, self.d).contiguous()
        y_new = y_new.to(device=dev, dtype=torch.float64).reshape(-1).contiguous()
        q, N, Np = X_new.shape[0], self.N, self.Np
        assert y_new.shape[0] == q and q > 0
        Np2 = max(Np, (N + q + 127) // 128 * 128)
        if Np2 == Np:
            alpha, linv, linv_t = self.alpha.clone(), self.linv.clone(), self.linv_t.clone()
        else:
            alpha = torch.zeros(Np2, dtype=torch.float64, device=dev)
            linv = torch.zeros(Np2, Np2, dtype=torch.float64, device=dev)
            linv_t = torch.zeros(Np2, Np2, dtype=torch.float64, device=dev)
            alpha[:Np].copy_(self.alpha)
            linv[:Np, :Np].copy_(self.linv)
            linv_t[:Np, :Np].copy_(self.linv_t)
        Xbuf = torch.empty(N + q, self.d, dtype=torch.float64, device=dev)
        Xbuf[:N].copy_(self.X)
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        nbytes = lib.cbo_model_append_workspace_bytes(Np2)
        ws = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev)
        st = lib.cbo_model_append_f64(Xbuf.data_ptr(), N, self.d, Np2, self.inv_lengthscale.data_ptr(), self.mean,
                                      self.outputscale, self.noise, alpha.data_ptr(), linv.data_ptr(), linv_t.data_ptr(),
                                      X_new.data_ptr(), y_new.data_ptr(), q, info.data_ptr(), ws.data_ptr(), nbytes,
                                      torch.cuda.current_stream().cuda_stream)
        _lib.check(st, "cbo_model_append_f64")
        if check_pd and int(info.item()) != 0:
            raise _lib.CboError(f"appended point {int(info.item()) - 1 - N} makes K + noise I numerically singular")
        return GPModelState(X=Xbuf, y=torch.cat([self.y, y_new]), mean=self.mean, lengthscale=self.lengthscale,
                            outputscale=self.outputscale, noise=self.noise, alpha=alpha, linv=linv, linv_t=linv_t,
                            inv_lengthscale=self.inv_lengthscale)

    def c_struct(self) -> "_lib.GpF64":
        return _lib.GpF64(self.N, self.d, self.Np, self.X.data_ptr(), self.alpha.data_ptr(), self.linv.data_ptr(),
                          self.linv_t.data_ptr(), self.inv_lengthscale.data_ptr(), self.mean, self.outputscale,
                          self.noise)

    def to(self, device) -> "GPModelState":
        device = torch.device(device)
        kw = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()}
        return GPModelState(**kw)
