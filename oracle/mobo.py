"""cEHVI oracle (test infrastructure only; "parity unpinned").

qConformalExpectedHypervolumeImprovement (conformalbo/mobo_acquisitions.py:24-45) for q = 1 and m independent
single-output GPs.  The reference's own `_conformal_integration` collapses the wrong axes for EHVI-shaped values and
raises unless b in {1, G} (verified there), and its multi-output label layout is scrambled by a
`.view` (App. C-6); like the product, this oracle implements the evident intent: per-output full-conformal masks,
product over outputs, sum of log-densities, botorch `_compute_qehvi` values (App. B-9), integration over the grid.
`impl='dense'` takes every per-output quantity from oracle/dense.py, `impl='closed'` from oracle/closed_form.py.
"""
from __future__ import annotations

import torch

from . import closed_form, dense


def box_decomposition_2d(ref_point: torch.Tensor, Y: torch.Tensor):
    """Non-dominated cells (maximisation, m = 2) of the Pareto front of Y above ref_point: (lower, upper) (n+1, 2)."""
    Y = Y[(Y > ref_point).all(-1)]
    order = torch.argsort(Y[:, 0])
    Y = Y[order]
    front = []
    best2 = -float("inf")
    for i in range(Y.shape[0] - 1, -1, -1):           # scan from largest y1: keep strictly improving y2
        if Y[i, 1] > best2:
            front.append(Y[i])
            best2 = float(Y[i, 1])
    front = torch.stack(front[::-1]) if front else Y.new_zeros(0, 2)   # y1 ascending, y2 descending
    n = front.shape[0]
    inf = float("inf")
    lower, upper = [], []
    for i in range(n + 1):
        l1 = float(ref_point[0]) if i == 0 else float(front[i - 1, 0])
        u1 = inf if i == n else float(front[i, 0])
        l2 = float(ref_point[1]) if i == n else float(front[i, 1])
        lower.append([l1, l2])
        upper.append([u1, inf])
    return torch.tensor(lower, dtype=torch.float64), torch.tensor(upper, dtype=torch.float64)


def cehvi(gps, Xq, *, alpha, temp, grid_res, eps_grid, z, eps_mc, cell_lower, cell_upper, ratio=None, randomized=False,
          u_rand=None, optimistic=False, impl="closed"):
    """gps: list of m GPState; Xq (b,1,d); eps_grid (G,m); eps_mc (S,G,m); u_rand (b,G,m).  Returns acq (b,)."""
    m = len(gps)
    b, G = Xq.shape[0], grid_res
    mask = torch.ones(b, G, dtype=torch.float64)
    opt = torch.ones(b, G, dtype=torch.float64)
    logp = torch.zeros(b, G, dtype=torch.float64)
    ood = torch.ones(b, dtype=torch.float64)
    mean_f, sd_f = [], []
    for k, gp in enumerate(gps):
        eg = eps_grid[:, k:k + 1]
        ur = None if u_rand is None else u_rand[:, :, k:k + 1]
        if impl == "closed":
            o = closed_form.conformal_acquisition(gp, Xq, kind="ucb", alpha=alpha, temp=temp, grid_res=G, eps_grid=eg, z=z,
                                                  eps_mc=torch.zeros(1, G, 1, dtype=torch.float64), ratio=ratio,
                                                  randomized=randomized, u_rand=ur, optimistic=optimistic, beta_prime=1.0)
            grid, lp, mk, od, mu, s = o["target_grid"][..., 0], o["logp"], o["mask"][..., 0], o["ood"][:, 0, 0], o["mu"][:, 0], o["S"][:, 0, 0]
            mf = mu[:, None] + (grid - mu[:, None]) * (1 - gp.noise / s)[:, None]
            sd = (gp.noise * (1 - gp.noise / s)).clamp_min(0).sqrt()
        else:
            grid3, lp, mk3, refit, od3 = dense.construct_conformal_bands(gp, Xq, alpha, temp, G, 0, eg, z, ratio, randomized, ur)
            grid, mk, od = grid3[..., 0], mk3[..., 0], od3[:, 0, 0]
            mu, _ = dense.posterior(gp, Xq, observation_noise=False)
            mu = mu[:, 0]
            mfull, cov = dense.refit_posterior_at_candidates(gp, refit, 1)
            mf, sd = mfull[..., 0], cov[:, 0, 0, 0].clamp_min(0).sqrt()     # the refit variance does not depend on g
        mask = mask * mk
        logp = logp + lp
        ood = ood * od
        if optimistic:
            opt = opt * (grid >= mu[:, None]).double()
        mean_f.append(mf)
        sd_f.append(sd)
    mean_f = torch.stack(mean_f, -1)                                   # (b,G,m)
    sd_f = torch.stack(sd_f, -1)                                       # (b,m)
    f = mean_f.unsqueeze(0) + sd_f[None, :, None, :] * eps_mc[:, None]  # (S,b,G,m)
    vol = (torch.minimum(f.unsqueeze(-2), cell_upper) - cell_lower).clamp_min(0).prod(-1)   # (S,b,G,cells)
    values = vol.sum(-1).mean(0)                                       # (b,G): _compute_qehvi takes the MC mean
    nonconf = (1 - mask) * opt
    nonconf = nonconf / nonconf.sum(1, keepdim=True).clamp_min(1e-6)
    conf = mask * opt / logp.exp().clamp_min(1e-6)
    conf = conf / conf.sum(1, keepdim=True).clamp_min(1e-6)
    return ((1 - ood)[:, None] * ((1 - alpha) * conf + alpha * nonconf) * values).sum(1)
