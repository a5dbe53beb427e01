"""cEHVI oracle (test infrastructure only).  Pinning: the inner expectation (`values`, b x G), the joint grid log-density and
the optimistic mask match what the reference's own `_conformal_fwd` hands to `_conformal_integration` (executed over
oracle/_shim: tests/golden/ref_path_ehvi.pt, tests/test_oracle_ref_path.py); the integration over the grid, the per-output
("intent") label layout of the masks and cNEHVI's baseline box decompositions are "parity unpinned" (see below).

qConformalExpectedHypervolumeImprovement (conformalbo/mobo_acquisitions.py:24-45) and
qConformalNoisyExpectedHypervolumeImprovement (:184-219) for joint batches of q points and m independent single-output GPs.  The reference's own `_conformal_integration` collapses the wrong axes for EHVI-shaped values and
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


def qehvi(f, cell_lower, cell_upper):
    """botorch `_compute_qehvi` before the MC mean.  f: (S, ..., q, m); cells (n_cells, m) or (S, n_cells, m)
    (one box decomposition per MC sample: qNoisyExpectedHypervolumeImprovement).  Returns (S, ...)."""
    from itertools import combinations
    S, q = f.shape[0], f.shape[-2]
    extra = f.dim() - 3
    if cell_lower.dim() == 2:
        shape = (1,) + (1,) * extra + cell_lower.shape
    else:
        shape = (S,) + (1,) * extra + cell_lower.shape[1:]
    lo, up = cell_lower.reshape(shape), cell_upper.reshape(shape)           # (S|1, ..., cells, m)
    total = torch.zeros(f.shape[:-2], dtype=f.dtype)
    for i in range(1, q + 1):
        for sub in combinations(range(q), i):
            v = f[..., list(sub), :].min(dim=-2).values                     # (S, ..., m)
            vol = (torch.minimum(v.unsqueeze(-2), up) - lo).clamp_min(0).prod(-1).sum(-1)
            total = total + (-1) ** (i + 1) * vol
    return total


def cehvi(gps, Xq, *, alpha, temp, grid_res, eps_grid, z, eps_mc, cell_lower, cell_upper, ratio=None, randomized=False,
          u_rand=None, optimistic=False, impl="closed", max_grid_refinements=0, X_baseline=None, lazy_prior_grad=False):
    """gps: list of m GPState; Xq (b,q,d); eps_grid (G,q,m); eps_mc (S,G,q,m) -- or (S,G,nb+q,m) with X_baseline (cNEHVI: the
    samples of f(X) come from the joint refit posterior over [X_baseline; X], cells are then (S,n_cells,m)); u_rand (b,G,q,m).
    Returns dict(acq (b,), values (b,G)).  Differentiable in Xq (torch autograd)."""
    mod = closed_form if impl == "closed" else dense
    b, q, _ = Xq.shape
    G = grid_res
    mask = torch.ones(b, G, dtype=torch.float64)
    opt = torch.ones(b, G, dtype=torch.float64)
    logp = torch.zeros(b, G, dtype=torch.float64)
    ood = torch.ones(b, dtype=torch.float64)
    fs = []
    for k, gp in enumerate(gps):
        o = mod.conformal_acquisition(gp, Xq, kind="raw" if X_baseline is None else "joint", alpha=alpha, temp=temp, grid_res=G,
                                      max_grid_refinements=max_grid_refinements, eps_grid=eps_grid[..., k], z=z,
                                      eps_mc=eps_mc[..., k], ratio=ratio, randomized=randomized,
                                      u_rand=None if u_rand is None else u_rand[..., k], optimistic=optimistic,
                                      beta_prime=1.0, X_baseline=X_baseline, lazy_prior_grad=lazy_prior_grad)
        mask = mask * o["mask"].prod(-1)                                   # product over the q points and over outputs
        logp = logp + o["logp"]
        ood = ood * o["ood"][:, 0, :].prod(-1)
        if optimistic:
            with torch.no_grad():
                mu = gp.mean + gp.k(Xq, gp.X) @ gp.alpha                   # (b,q)
            opt = opt * (o["target_grid"] >= mu.unsqueeze(1)).double().prod(-1)
        fs.append(o["values"])                                             # (S,b,G,q)
    f = torch.stack(fs, -1)                                                # (S,b,G,q,m)
    values = qehvi(f, cell_lower, cell_upper).mean(0)                      # (b,G)
    nonconf = (1 - mask) * opt
    nonconf = nonconf / nonconf.sum(1, keepdim=True).clamp_min(1e-6)
    conf = mask * opt / logp.exp().clamp_min(1e-6)
    conf = conf / conf.sum(1, keepdim=True).clamp_min(1e-6)
    acq = ((1 - ood)[:, None] * ((1 - alpha) * conf + alpha * nonconf) * values).sum(1)
    return dict(acq=acq, values=values, mask=mask, logp=logp)


def nehvi_cells(gps, X_baseline, eps_b, ref_point):
    """qNoisyExpectedHypervolumeImprovement._set_cell_bounds (botorch 0.7, from memory -- "parity unpinned"): sample the
    baseline values under the model, f_b^s = mean + chol(cov) eps_b[s] per output (psd-safe Cholesky: escalating jitter), and
    decompose the non-dominated region of every sample.  eps_b (S, nb, m) -> (lower, upper) (S, n_cells, m), zero-volume padding."""
    S = eps_b.shape[0]
    fb = []
    for k, gp in enumerate(gps):
        Kb = gp.k(X_baseline, gp.X)
        mu = gp.mean + Kb @ gp.alpha
        Vb = torch.linalg.solve_triangular(gp.L, Kb.T, upper=False)
        cov = gp.k(X_baseline, X_baseline) - Vb.T @ Vb
        L, info = torch.linalg.cholesky_ex(cov)
        jitter = 1e-8
        while int(info) != 0:
            L, info = torch.linalg.cholesky_ex(cov + jitter * torch.eye(cov.shape[0], dtype=cov.dtype))
            jitter *= 10.0
        fb.append(mu.unsqueeze(0) + eps_b[:, :, k] @ L.T)
    fb = torch.stack(fb, -1)                                               # (S, nb, m)
    cells = [box_decomposition_2d(ref_point, fb[s]) for s in range(S)]
    n = max(c[0].shape[0] for c in cells)
    pad = ref_point.reshape(1, -1).double()
    lo = torch.stack([torch.cat([c[0], pad.expand(n - c[0].shape[0], -1)]) for c in cells])
    up = torch.stack([torch.cat([c[1], pad.expand(n - c[1].shape[0], -1)]) for c in cells])
    return lo, up, fb
