"""Exact-GP building blocks of the oracle (plain torch fp64 on CPU; test infrastructure only).

Restates the slice of BoTorch 0.7 `SingleTaskGP` / GPyTorch 1.9.0 that the reference path uses
(/root/reference/conformalbo/utils.py:382-390 builds `SingleTaskGP(train_X, train_Y)`):
ConstantMean + ScaleKernel(MaternKernel(nu=2.5, ARD)) + homoskedastic Gaussian noise.
Third-party semantics are restated from their published definitions.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

SQRT5 = math.sqrt(5.0)


def matern52(x1: torch.Tensor, x2: torch.Tensor, lengthscale: torch.Tensor, outputscale) -> torch.Tensor:
    """outputscale * (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r), r = ||(x1-x2)/lengthscale||.

    GPyTorch MaternKernel(nu=2.5) under ScaleKernel (App. B-1).  Distances are formed from
    explicit differences (GPyTorch uses the mean-centred GEMM identity; both are exact up to
    rounding, the explicit form is the better-conditioned one for an oracle).
    x1: (..., n1, d), x2: (..., n2, d) -> (..., n1, n2)
    """
    sq = None
    for k in range(x1.shape[-1]):   # accumulate over d: avoids an (n1, n2, d) temporary
        dk = (x1[..., :, None, k] - x2[..., None, :, k]) / lengthscale[k]
        sq = dk * dk if sq is None else sq + dk * dk
    r = sq.clamp_min(1e-30).sqrt()
    # r == 0 exactly on the diagonal of k(X, X): clamp gives 1e-15, value 1 to rounding.
    return outputscale * (1.0 + SQRT5 * r + (5.0 / 3.0) * r * r) * torch.exp(-SQRT5 * r)


@dataclass
class GPState:
    """One output's exact GP.  A = K + noise*I = L L^T, alpha = A^{-1}(y - mean)."""
    X: torch.Tensor            # (N, d)
    y: torch.Tensor            # (N,)
    mean: float
    lengthscale: torch.Tensor  # (d,)
    outputscale: float
    noise: float
    L: torch.Tensor            # (N, N) lower Cholesky of A
    alpha: torch.Tensor        # (N,)

    @property
    def N(self):
        return self.X.shape[0]

    def k(self, a, b):
        return matern52(a, b, self.lengthscale, self.outputscale)


def build_state(X, y, mean, lengthscale, outputscale, noise) -> GPState:
    X = X.double()
    y = y.double().reshape(-1)
    lengthscale = torch.as_tensor(lengthscale, dtype=torch.float64).reshape(-1)
    K = matern52(X, X, lengthscale, outputscale)
    A = K + noise * torch.eye(X.shape[0], dtype=torch.float64)
    L = torch.linalg.cholesky(A)
    alpha = torch.cholesky_solve((y - mean).unsqueeze(-1), L).squeeze(-1)
    return GPState(X, y, float(mean), lengthscale, float(outputscale), float(noise), L, alpha)


@dataclass
class RatioMLP:
    """Weights of RatioEstimator._target_network (ratio_estimation.py:73-86): d->64->ReLU->64->ReLU->1."""
    W1: torch.Tensor  # (64, d)
    b1: torch.Tensor  # (64,)
    W2: torch.Tensor  # (64, 64)
    b2: torch.Tensor  # (64,)
    W3: torch.Tensor  # (1, 64)
    b3: torch.Tensor  # (1,)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        """RatioEstimator.forward, ratio_estimation.py:97-99."""
        h = torch.relu(x @ self.W1.T + self.b1)
        h = torch.relu(h @ self.W2.T + self.b2)
        logit = (h @ self.W3.T + self.b3).squeeze(-1)
        p = torch.sigmoid(logit)
        return p.clamp_max(1 - 1e-6) / (1 - p).clamp_min(1e-6)


def make_ratio_mlp(d: int, seed: int = 1, hidden: int = 64) -> RatioMLP:
    """torch default nn.Linear init under a fixed seed, copied to the target net."""
    g = torch.Generator().manual_seed(seed)

    def lin(fan_in, fan_out):
        bound = 1.0 / math.sqrt(fan_in)
        W = (torch.rand(fan_out, fan_in, generator=g, dtype=torch.float64) * 2 - 1) * bound
        b = (torch.rand(fan_out, generator=g, dtype=torch.float64) * 2 - 1) * bound
        return W, b

    W1, b1 = lin(d, hidden)
    W2, b2 = lin(hidden, hidden)
    W3, b3 = lin(hidden, 1)
    return RatioMLP(W1, b1, W2, b2, W3, b3)
