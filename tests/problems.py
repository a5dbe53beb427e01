"""Seeded synthetic problems shared by the CPU and GPU tests."""
import math

import torch

from oracle import gp as ogp


def ackley(X):
    """-Ackley_d on [-32.768, 32.768]^d from the unit cube (botorch test function, utils.py:61-115)."""
    x = -32.768 + 65.536 * X
    d = x.shape[-1]
    a, b, c = 20.0, 0.2, 2 * math.pi
    f = -a * torch.exp(-b * torch.sqrt((x * x).mean(-1))) - torch.exp(torch.cos(c * x).mean(-1)) + a + math.e
    return -f


def make_problem(N, d, seed=0, noise=1e-2, lengthscale=0.5, outputscale=1.0, mean=0.0, fn="ackley"):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(N, d, generator=g, dtype=torch.float64)
    if fn == "ackley":
        f = ackley(X)
    else:
        f = torch.sin(3.0 * X.sum(-1))
    y = f + 0.1 * f.std() * torch.randn(N, generator=g, dtype=torch.float64)
    y = (y - y.mean()) / y.std()
    ls = torch.full((d,), lengthscale, dtype=torch.float64) * (1.0 + 0.1 * torch.arange(d, dtype=torch.float64) / max(d, 1))
    return dict(X=X, y=y, mean=mean, lengthscale=ls, outputscale=outputscale, noise=noise)


def make_inputs(N, d, b, q, G, S=1, Sp=8, seed=0):
    g = torch.Generator().manual_seed(1000 + seed)
    Xq = torch.rand(b, q, d, generator=g, dtype=torch.float64)
    eps_grid = torch.randn(G, q, generator=g, dtype=torch.float64)
    z = 2.0 * ((torch.rand(N + q, Sp, generator=g) > 0.5).double() - 0.5)
    eps_mc = torch.randn(S, G, q, generator=g, dtype=torch.float64)
    u_rand = torch.rand(b, G, q, generator=g, dtype=torch.float64)
    return dict(Xq=Xq, eps_grid=eps_grid, z=z, eps_mc=eps_mc, u_rand=u_rand)


def oracle_state(p):
    return ogp.build_state(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"])
