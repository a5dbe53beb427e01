"""Seeded synthetic workloads of BASELINE.json / SURVEY.md §8d (plain CPU torch tensors; no arithmetic of the path)."""
from __future__ import annotations

import math

import torch


def ackley(X: torch.Tensor) -> torch.Tensor:
    """-Ackley_d evaluated on the unit cube mapped to [-32.768, 32.768]^d (task of BASELINE config 2)."""
    x = -32.768 + 65.536 * X
    a, b, c = 20.0, 0.2, 2 * math.pi
    f = -a * torch.exp(-b * torch.sqrt((x * x).mean(-1))) - torch.exp(torch.cos(c * x).mean(-1)) + a + math.e
    return -f


def make_problem(N, d, seed=0, noise=1e-2, lengthscale=0.5, outputscale=1.0, mean=0.0, fn="ackley", ard_spread=0.0):
    """Training data + FIXED hyper-parameters (no fitting, for determinism)."""
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(N, d, generator=g, dtype=torch.float64)
    f = ackley(X) if fn == "ackley" else torch.sin(3.0 * X.sum(-1))
    y = f + 0.1 * f.std() * torch.randn(N, generator=g, dtype=torch.float64)
    y = (y - y.mean()) / y.std()
    ls = torch.full((d,), lengthscale, dtype=torch.float64) * (1.0 + ard_spread * torch.arange(d, dtype=torch.float64) / max(d, 1))
    return dict(X=X, y=y, mean=mean, lengthscale=ls, outputscale=outputscale, noise=noise)


def make_inputs(N, d, b, q, G, S=1, Sp=8, seed=0):
    """Candidates and the injected random inputs."""
    g = torch.Generator().manual_seed(1000 + seed)
    Xq = torch.rand(b, q, d, generator=g, dtype=torch.float64)
    eps_grid = torch.randn(G, q, generator=g, dtype=torch.float64)
    z = 2.0 * ((torch.rand(N + q, Sp, generator=g) > 0.5).double() - 0.5)
    eps_mc = torch.randn(S, G, q, generator=g, dtype=torch.float64)
    u_rand = torch.rand(b, G, q, generator=g, dtype=torch.float64)
    return dict(Xq=Xq, eps_grid=eps_grid, z=z, eps_mc=eps_mc, u_rand=u_rand)


def make_ratio_weights(d: int, seed: int = 1, hidden: int = 64):
    """nn.Linear default init under a fixed seed (W1,b1,W2,b2,W3,b3) -- a stand-in for a trained target network."""
    g = torch.Generator().manual_seed(seed)

    def lin(fan_in, fan_out):
        bound = 1.0 / math.sqrt(fan_in)
        W = (torch.rand(fan_out, fan_in, generator=g, dtype=torch.float64) * 2 - 1) * bound
        bb = (torch.rand(fan_out, generator=g, dtype=torch.float64) * 2 - 1) * bound
        return W, bb

    W1, b1 = lin(d, hidden)
    W2, b2 = lin(hidden, hidden)
    W3, b3 = lin(hidden, 1)
    return W1, b1, W2, b2, W3, b3
