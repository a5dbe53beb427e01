"""Closed-form restatement of the full-conformal acquisition path (CPU, torch fp64).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  "parity unpinned" at the botorch/gpytorch
boundary; pinned by agreement with oracle/dense.py (tests/test_oracle_consistency.py).

Same inputs/outputs as oracle/dense.py but evaluated with the O(N^2)-per-candidate algebra of
SURVEY.md App. A (this is the algorithm the CUDA kernels implement, written independently of
them in plain torch so that it can serve as their checker and as the algorithm-matched CPU
baseline):
  A-1  u = k(X,x), v = L^-1 u, w = L^-T v, mu = m + u.alpha, S = K_qq + noise I - V^T V
  A-2  y_g = mu + chol(c S) eps_g,   logp_g
  A-3  B_g = S^-1 (y_g - mu),  residuals  r_i = noise (alpha_i - W_i. B_g),  r_c = noise B_g
  A-4  C = S^-1 (z_c - U^T T),  (A_new^-1 z) = [T - W C ; C],  var = noise (1 + (S_p - noise sum_s z (A_new^-1 z)) / (S_p + 1e-6))
  A-5  scores, hard count, soft rank, importance weights, randomised step, mask, ood
  A-6  refit posterior at the candidates: mean_f = mu + (I - noise S^-1)(y_g - mu), cov_f = noise (I - noise S^-1)
  A-7  importance-weighted integration over the grid
Reference lines: helpers.py:94-333,390-412; acquisitions.py:53-76,88-98,151-159,225-248.
"""
from __future__ import annotations

import math

import torch

from .gp import GPState

LOG2PI = math.log(2.0 * math.pi)


def candidate_solves(gp: GPState, Xq: torch.Tensor):
    """A-1 for a batch: returns U (b,q,N), W (b,q,N), mu (b,q), S (b,q,q)."""
    U = gp.k(Xq, gp.X)
    V = torch.linalg.solve_triangular(gp.L, U.transpose(-1, -2), upper=False)            # (b,N,q)
    W = torch.linalg.solve_triangular(gp.L.T, V, upper=True).transpose(-1, -2)           # (b,q,N)
    mu = gp.mean + U @ gp.alpha
    S = gp.k(Xq, Xq) + gp.noise * torch.eye(Xq.shape[-2], dtype=U.dtype) - V.transpose(-1, -2) @ V
    return U, W, mu, S


def conformal_acquisition(gp: GPState, Xq, *, kind, alpha, temp, grid_res, max_grid_refinements=0,
                          eps_grid, z, eps_mc, ratio=None, randomized=False, u_rand=None,
                          optimistic=True, beta_prime=None, best_f=None, X_baseline=None, lazy_prior_grad=False):
    """lazy_prior_grad: GRADIENT semantics only -- the prior block k(x_j, x_k) of the grid covariance tracks d/dX, as in the
    reference's lazily evaluated predictive covariance (see oracle/dense.py::construct_conformal_bands); default: the grid is a
    constant, the intent of the reference's no_grad block."""
    b, q, _ = Xq.shape
    N, G, noise = gp.N, grid_res, gp.noise
    Sp = z.shape[-1]
    U, W, mu, S = candidate_solves(gp, Xq)
    eye = torch.eye(q, dtype=U.dtype)
    Sinv = torch.linalg.inv(S)
    if eps_grid.ndim == 2:
        eps_grid = eps_grid.unsqueeze(0).expand(b, -1, -1)

    # ---- g-independent pieces -------------------------------------------------------------
    T = torch.cholesky_solve(z[:N], gp.L)                                   # (N,Sp) once per (model, z)
    C = Sinv @ (z[N:].unsqueeze(0) - U @ T)                                 # (b,q,Sp)
    Ainv_z_train = T.unsqueeze(0) - W.transpose(-1, -2) @ C                 # (b,N,Sp)
    denom = Sp + 1e-6
    var_train = noise * (1.0 + (Sp - noise * (z[:N] * Ainv_z_train).sum(-1)) / denom)     # (b,N)
    var_cand = noise * (1.0 + (Sp - noise * (z[N:] * C).sum(-1)) / denom)                 # (b,q)
    if ratio is None:
        dr_train = torch.ones(N, dtype=U.dtype)
        dr_cand = torch.ones(b, q, dtype=U.dtype)
    else:
        dr_train = ratio(gp.X).clamp_min(1e-6)
        dr_cand = ratio(Xq).clamp_min(1e-6)
    omega_c = dr_cand / (dr_train.sum() + dr_cand)                          # (b,q): own-column normalisation
    ood = torch.sigmoid((omega_c - alpha) / temp) if temp > 0 else (omega_c >= alpha).to(U)

    def one_pass(cov):
        # the grid and its log-density are built under torch.no_grad() in the reference (helpers.py:118-119): constants
        # w.r.t. the candidate; everything below (the refit on (x, y_g)) is differentiable in x
        with torch.set_grad_enabled(bool(lazy_prior_grad) and torch.is_grad_enabled()):
            Lc = torch.linalg.cholesky(cov)
            dev0 = (Lc.unsqueeze(1) @ eps_grid.unsqueeze(-1)).squeeze(-1)   # (b,G,q) = y_g - mu
            sol = torch.linalg.solve_triangular(Lc.unsqueeze(1), dev0.unsqueeze(-1), upper=False).squeeze(-1)
            logdet = 2.0 * torch.diagonal(Lc, dim1=-2, dim2=-1).log().sum(-1)
            logp = -0.5 * (sol.pow(2).sum(-1) + logdet.unsqueeze(1) + q * LOG2PI)
            grid_const = mu.detach().unsqueeze(1) + dev0
        dev = grid_const - mu.unsqueeze(1)                                  # same numbers; carries d/dx of mu
        Bg = dev @ Sinv.transpose(-1, -2)                                   # (b,G,q)  S^-1 (y_g - mu)
        r_train = noise * (gp.alpha.view(1, 1, N) - Bg @ W)                 # (b,G,N)
        r_cand = noise * Bg                                                 # (b,G,q)
        sc_train = -r_train.pow(2) / (2 * var_train.unsqueeze(1)) - 0.5 * var_train.log().unsqueeze(1) - 0.5 * LOG2PI
        sc_cand = -r_cand.pow(2) / (2 * var_cand.unsqueeze(1)) - 0.5 * var_cand.log().unsqueeze(1) - 0.5 * LOG2PI
        mask = torch.empty(b, G, q, dtype=U.dtype)
        total_dr = dr_train.sum()
        for j in range(q):
            thr = sc_cand[..., j]                                           # (b,G)
            ge_t = (thr.unsqueeze(-1) >= sc_train)                          # (b,G,N)
            ge_c = (thr.unsqueeze(-1) >= sc_cand)                           # (b,G,q) includes self
            n_g = ge_t.sum(-1) + ge_c.sum(-1)
            if temp > 0:
                rho_t = torch.sigmoid((thr.unsqueeze(-1) - sc_train) / temp)
                rho_c = torch.sigmoid((thr.unsqueeze(-1) - sc_cand) / temp)
                scale = n_g / (rho_t.sum(-1) + rho_c.sum(-1))
                rho_t, rho_self = rho_t * scale.unsqueeze(-1), rho_c[..., j] * scale
            else:
                rho_t, rho_self = ge_t.to(U), torch.ones(b, G, dtype=U.dtype)
            norm = total_dr + dr_cand[:, j]                                 # (b,)
            M = ((rho_t * dr_train).sum(-1) + rho_self * dr_cand[:, j].unsqueeze(-1)) / norm.unsqueeze(-1)
            if randomized:
                if j == q - 1:                                              # helpers.py:307 takes the LAST row
                    D = (rho_self * (dr_cand[:, j] / norm).unsqueeze(-1)).clamp_min(1e-6)
                else:
                    D = torch.full_like(M, 1e-6)
                p = ((M - alpha) / D).clamp(0.0, 1.0)
                use_ub = (u_rand[..., j] < p).to(U)
                M = M - (1.0 - use_ub) * D
            mask[..., j] = torch.sigmoid((M - alpha) / temp) if temp > 0 else (M > alpha).to(U)
        return grid_const, logp, mask, dev

    cov = S.detach().clone()
    if lazy_prior_grad and Xq.requires_grad:
        Kxx = gp.k(Xq, Xq)
        cov = cov + (Kxx - Kxx.detach())                                    # value unchanged, d/dX of the prior block only
    best = (float("inf"),)
    for _ in range(max_grid_refinements + 1):
        grid, logp, mask, dev = one_pass(cov)
        num_acc = (mask.detach() >= 0.5).float().sum(1)                     # float32 statistics, as helpers.py:150-167
        diff = num_acc / G - 0.5
        score = diff.pow(2).mean()
        if score < best[0]:
            best = (score, grid, logp, mask, dev)
        if num_acc.min() < 2 or num_acc.max() > G - 2:
            cov = (1.0 + diff.sign() * diff.abs().pow(2.0 / q)).unsqueeze(-1) * cov
        else:
            break
    _, grid, logp, mask, dev = best

    if kind in ("nei", "joint"):
        # joint posterior of the refit model over [X_baseline; X] (acquisitions.py:115-134).  With c = Cov_D(f(X_b), y(X))
        # (nb x q), Cov_D(f(X), y(X)) = S - noise I and B_g = S^-1 (y_g - mu):
        #   mean = [mu_b + c B_g ; mu + (I - noise S^-1)(y_g - mu)]
        #   cov  = [[Sigma_bb - c S^-1 c^T, noise c S^-1], [., noise (I - noise S^-1)]]      (g-independent)
        nb = X_baseline.shape[0]
        Kb = gp.k(X_baseline, gp.X)                                        # (nb,N)
        mu_b = gp.mean + Kb @ gp.alpha
        Vb = torch.linalg.solve_triangular(gp.L, Kb.T, upper=False)
        Sig_bb = gp.k(X_baseline, X_baseline) - Vb.T @ Vb
        c = gp.k(X_baseline.unsqueeze(0).expand(b, nb, -1), Xq) - Kb.unsqueeze(0) @ W.transpose(-1, -2)     # (b,nb,q)
        Sinv_s = 0.5 * (Sinv + Sinv.transpose(-1, -2))
        cSi = c @ Sinv_s                                                   # (b,nb,q)
        cov = torch.zeros(b, nb + q, nb + q, dtype=U.dtype)
        cov[:, :nb, :nb] = Sig_bb - cSi @ c.transpose(-1, -2)
        cov[:, :nb, nb:] = noise * cSi
        cov[:, nb:, :nb] = cov[:, :nb, nb:].transpose(-1, -2)
        cov[:, nb:, nb:] = noise * (eye - noise * Sinv_s)
        cov = 0.5 * (cov + cov.transpose(-1, -2))
        root = torch.linalg.cholesky(cov)
        Bg_v = dev @ Sinv.transpose(-1, -2)                                # (b,G,q)
        mean = torch.cat([mu_b.view(1, 1, nb) + Bg_v @ c.transpose(-1, -2),
                          mu.unsqueeze(1) + dev @ (eye - noise * Sinv).transpose(-1, -2)], dim=-1)          # (b,G,nb+q)
        samples = mean.unsqueeze(0) + torch.einsum("bij,sgj->sbgi", root, eps_mc)
        fb, fx = samples[..., :nb], samples[..., nb:]                      # (S,b,G,nb), (S,b,G,q)
        if kind == "joint":                                                # cNEHVI: the samples of f(X) themselves
            values = fx
        elif temp > 0:
            diffs = (fx - temp * torch.logsumexp(fb / temp, -1, keepdim=True)) / temp
            values = temp * (torch.clamp_min(diffs, 0) + torch.log1p(torch.exp(-diffs.abs())))
        else:
            values = torch.clamp_min(fx - fb.max(-1, keepdim=True).values, 0.0)
    # ---- A-6 values under the refit model ------------------------------------------------
    shrink = eye - noise * Sinv                                            # (b,q,q)
    mean_f = mu.unsqueeze(1) + dev @ shrink.transpose(-1, -2)              # (b,G,q)
    cov_f = noise * shrink
    cov_f = 0.5 * (cov_f + cov_f.transpose(-1, -2))
    root = cov_f.clamp_min(0.0).sqrt() if q == 1 else torch.linalg.cholesky(cov_f)
    if kind not in ("nei", "joint"):
        samples = mean_f.unsqueeze(0) + torch.einsum("bij,sgj->sbgi", root, eps_mc)
    if kind == "ucb":
        m = samples.mean(0)
        values = m + beta_prime * (samples - m).abs()
    elif kind == "ei":
        x = (samples - best_f) / temp
        values = temp * (torch.clamp_min(x, 0) + torch.log1p(torch.exp(-x.abs())))
    elif kind == "raw":                                                    # cEHVI: the samples of f(X) themselves
        values = samples
    elif kind in ("nei", "joint"):
        pass                                                               # computed above from the joint posterior
    else:
        raise ValueError(kind)

    # ---- A-7 integration -----------------------------------------------------------------
    opt = (grid >= mu.detach().unsqueeze(1)).to(U) if optimistic else torch.ones_like(grid)
    nonconf = (1.0 - mask) * opt
    nonconf = nonconf / nonconf.sum(1, keepdim=True).clamp_min(1e-6)
    conf = mask * opt / logp.exp().clamp_min(1e-6).unsqueeze(-1)
    conf = conf / conf.sum(1, keepdim=True).clamp_min(1e-6)
    wts = (1.0 - ood).unsqueeze(1) * ((1.0 - alpha) * conf + alpha * nonconf)     # (b,G,q)
    res = (wts.unsqueeze(0) * values).sum(2)                                      # (S,b,q)
    acq = res.max(-1).values.mean(0)
    return dict(acq=acq, target_grid=grid, logp=logp, mask=mask, ood=ood.unsqueeze(1).expand(b, G, q),
                values=values, mu=mu, S=S, var_train=var_train, var_cand=var_cand)
