"""Dense literal restatement of the reference's full-conformal acquisition path (CPU, torch fp64).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  "parity unpinned": the reference holds no
tests or golden vectors for these functions and cannot be executed here (botorch/gpytorch absent);
this file restates its algorithm op-by-op and is cross-checked against oracle/closed_form.py.

It deliberately does what the reference does and the product does NOT do: for every
(candidate, grid point) pair it materialises the dense (N+q)x(N+q) system of the refit GP,
factorises it, and evaluates means / Hutchinson variances / acquisition samples from that dense
factor -- O(b*G*(N+q)^3) flops, O(b*G*(N+q)^2) memory.

Follows (file:line in /root/reference):
  construct_conformal_bands    conformalbo/helpers.py:94-202
  refine_grid_dist             conformalbo/helpers.py:205-215
  conformal_gp_regression      conformalbo/helpers.py:218-333
  est_train_post_var           conformalbo/helpers.py:390-412
  ConformalAcquisition         conformalbo/acquisitions.py:29-76
  qConformalExpectedImprovement._nonconformal_fwd      conformalbo/acquisitions.py:88-98
  qConformalUpperConfidenceBound._nonconformal_fwd     conformalbo/acquisitions.py:151-159
Third-party semantics (GP posterior, fantasy refit = exact GP on the augmented data, sampler base
sample layout, MVN log_prob / rsample) restated per SURVEY.md App. B.  Randomness is injected:
  eps_grid (G,q) or (b,G,q)   base samples of the grid sampler        (helpers.py:132-133)
  z        (N+q, S_p) of +-1  Hutchinson probes                       (helpers.py:399-400)
  eps_mc   (S,G,q)            base samples of the acquisition sampler (acquisitions.py:155)
  u_rand   (b,G,q) in [0,1)   uniforms behind Bernoulli.sample        (helpers.py:311)
"""
from __future__ import annotations

import math

import torch

from .gp import GPState
from .pure import conformal_integration, softplus

LOG2PI = math.log(2.0 * math.pi)


def _chol(M: torch.Tensor) -> torch.Tensor:
    """psd_safe_cholesky: plain Cholesky of the lower triangle, jitter only on failure (App. B-3)."""
    L, info = torch.linalg.cholesky_ex(M)
    if not bool((info != 0).any()):
        return L
    for jitter in (1e-8, 1e-7, 1e-6, 1e-5, 1e-4):
        L, info = torch.linalg.cholesky_ex(M + jitter * torch.eye(M.shape[-1], dtype=M.dtype))
        if not bool((info != 0).any()):
            return L
    raise RuntimeError("matrix not positive definite")


def posterior(gp: GPState, Xq: torch.Tensor, observation_noise: bool):
    """model.posterior(X, observation_noise): mean (b,q), covariance (b,q,q).  helpers.py:119."""
    U = gp.k(Xq, gp.X)                                   # (b,q,N)
    mean = gp.mean + U @ gp.alpha                        # (b,q)
    V = torch.linalg.solve_triangular(gp.L, U.transpose(-1, -2), upper=False)   # (b,N,q)
    cov = gp.k(Xq, Xq) - V.transpose(-1, -2) @ V
    if observation_noise:
        cov = cov + gp.noise * torch.eye(Xq.shape[-2], dtype=cov.dtype)
    return mean, cov


def conformal_gp_regression(gp: GPState, Xq, target_grid, alpha, temp, z, ratio=None,
                            randomized=False, u_rand=None, score_labels=None):
    """helpers.py:218-333 for one output (t=1).  Xq (b,q,d), target_grid (b,G,q).

    Returns conf_pred_mask (b,G,q), refit (dict of dense per-(b,G) quantities), q_conf_scores
    (b,G,q), ood_mask (b,G,q).
    """
    b, q, d = Xq.shape
    G = target_grid.shape[1]
    N = gp.N
    n = N + q
    # helpers.py:239-243  condition_on_observations == exact GP on the augmented data set
    X_full = torch.cat([gp.X.expand(b, G, N, d), Xq.unsqueeze(1).expand(b, G, q, d)], dim=-2)  # (b,G,n,d)
    labels = torch.cat([gp.y.expand(b, G, N), target_grid], dim=-1)                             # (b,G,n)
    K_full = gp.k(X_full, X_full)                                                               # (b,G,n,n)
    A_full = K_full + gp.noise * torch.eye(n, dtype=K_full.dtype)
    L_full = _chol(A_full)                                                                      # fresh potrf per (b,G)
    alpha_new = torch.cholesky_solve((labels - gp.mean).unsqueeze(-1), L_full).squeeze(-1)      # (b,G,n)

    # helpers.py:266  posterior mean of the refit model at its own training inputs
    mean_train = gp.mean + (K_full @ alpha_new.unsqueeze(-1)).squeeze(-1)                       # (b,G,n)
    # helpers.py:267 -> 390-412  Hutchinson diagonal estimate, observation_noise=True
    denom = (z * z).sum(-1) + 1e-6                                                              # (n,)
    Ainv_z = torch.cholesky_solve(z.expand(b, G, n, z.shape[-1]), L_full)
    approx_eigs = (z * (K_full @ Ainv_z)).sum(-1) / denom + 1.0
    post_var = gp.noise * approx_eigs                                                           # (b,G,n)

    # helpers.py:268-272  scores = log N(label; mean, var)
    std = post_var.sqrt()
    # score_labels: the labels the scores are evaluated AT when they are not the labels the model was refit on -- the
    # multi-output `.view` scramble of helpers.py:257-258 (oracle/multi.py)
    lab = labels if score_labels is None else score_labels
    conf_scores = -((lab - mean_train) ** 2) / (2 * post_var) - std.log() - 0.5 * LOG2PI        # (b,G,n)

    threshold = conf_scores[..., N:]                                                            # (b,G,q)
    rank_mask = (threshold.unsqueeze(-2) >= conf_scores.unsqueeze(-1)).to(conf_scores)          # (b,G,n,q)
    if temp > 0:
        mask_sum = rank_mask.sum(dim=-2, keepdim=True)
        rank_mask = 1.0 - torch.sigmoid((conf_scores.unsqueeze(-1) - threshold.unsqueeze(-2)) / temp)
        rank_mask = rank_mask * (mask_sum / rank_mask.sum(dim=-2, keepdim=True))

    # helpers.py:285-298  importance weights
    imp_weights = torch.zeros_like(rank_mask)
    idx = torch.arange(q)
    if ratio is None:
        imp_weights[..., :N, :] = 1.0
        imp_weights[..., N + idx, idx] = 1.0
    else:
        dr_old = ratio(X_full[..., :N, :]).clamp_min(1e-6)          # (b,G,N)  (recomputed per (b,G) as in the reference)
        dr_new = ratio(X_full[..., N:, :]).clamp_min(1e-6)          # (b,G,q)
        imp_weights[..., :N, :] = dr_old.unsqueeze(-1)
        imp_weights[..., N + idx, idx] = dr_new
    imp_weights = imp_weights / imp_weights.sum(dim=-2, keepdim=True)

    masked_weights = rank_mask * imp_weights
    cum_weights = masked_weights.cumsum(-2)
    if randomized:
        diff = masked_weights[..., -1, :].clamp_min(1e-6)
        ub_prob = ((cum_weights[..., -1, :] - alpha) / diff).clamp(0.0, 1.0)
        use_ub = (u_rand < ub_prob).to(diff)                        # Bernoulli(p).sample() == [u < p]
        total = cum_weights[..., -1, :] - (1.0 - use_ub) * diff
    else:
        total = cum_weights[..., -1, :]                             # (b,G,q)

    if temp > 0:
        conf_pred_mask = torch.sigmoid((total - alpha) / temp)
        ood_mask = torch.sigmoid((imp_weights[..., N:, :].sum(-2) - alpha) / temp)
    else:
        conf_pred_mask = (total > alpha).to(total)
        ood_mask = (imp_weights[..., N:, :].sum(-2) >= alpha).to(total)

    refit = dict(K_full=K_full, L_full=L_full, alpha_new=alpha_new)
    return conf_pred_mask, refit, threshold, ood_mask


def refit_posterior_at_candidates(gp: GPState, refit, q: int):
    """conditioned_model.posterior(X_exp) (no observation noise): mean (b,G,q), cov (b,G,q,q)."""
    K_full, L_full, alpha_new = refit["K_full"], refit["L_full"], refit["alpha_new"]
    N = gp.N
    kq = K_full[..., N:, :]                                          # (b,G,q,n)
    mean_f = gp.mean + (kq @ alpha_new.unsqueeze(-1)).squeeze(-1)
    sol = torch.cholesky_solve(kq.transpose(-1, -2), L_full)         # (b,G,n,q)
    cov_f = K_full[..., N:, N:] - kq @ sol
    return mean_f, cov_f


def nei_values(gp, refit, Xq, X_baseline, eps_mc, temp, raw=False):
    """qConformalNoisyExpectedImprovement._nonconformal_fwd (acquisitions.py:115-134): joint posterior of the refit model over
    X_full = [X_baseline; X], joint samples (base samples (S,G,nb+q)), smooth improvement over the baseline."""
    b, q, d = Xq.shape
    N, nb = gp.N, X_baseline.shape[0]
    K_full, L_full, alpha_new = refit["K_full"], refit["L_full"], refit["alpha_new"]
    G = K_full.shape[1]
    X_train_full = torch.cat([gp.X.expand(b, N, d), Xq], dim=1)                                  # (b,N+q,d)
    X_eval = torch.cat([X_baseline.expand(b, nb, d), Xq], dim=1)                                 # (b,nb+q,d)
    Kx = gp.k(X_eval, X_train_full).unsqueeze(1).expand(b, G, nb + q, N + q)                     # cross-covariance (same for all g)
    mean = gp.mean + (Kx @ alpha_new.unsqueeze(-1)).squeeze(-1)                                  # (b,G,nb+q)
    sol = torch.cholesky_solve(Kx.transpose(-1, -2), L_full)
    cov = gp.k(X_eval, X_eval).unsqueeze(1) - Kx @ sol                                           # (b,G,nb+q,nb+q)
    cov = 0.5 * (cov + cov.transpose(-1, -2))
    root = _chol(cov)
    samples = mean.unsqueeze(0) + (root.unsqueeze(0) @ eps_mc.unsqueeze(1).unsqueeze(-1)).squeeze(-1)   # (S,b,G,nb+q)
    if raw:                                                          # cNEHVI (mobo_acquisitions.py:204-211): f(X) of the joint sample
        return samples[..., -q:]
    diffs = samples[..., -q:] - softplus(samples[..., :-q], -1, temp, keepdim=True)
    return softplus(torch.stack([diffs, torch.zeros_like(diffs)], dim=-1), -1, temp)             # (S,b,G,q)


def nonconformal_values(gp, refit, q, eps_mc, kind, *, beta_prime=None, best_f=None, temp=None):
    """_nonconformal_fwd of cUCB (acquisitions.py:151-159) / cEI (acquisitions.py:88-98).

    eps_mc (S,G,q) -> values (S,b,G,q).  Identity objective, no posterior transform.
    """
    mean_f, cov_f = refit_posterior_at_candidates(gp, refit, q)
    cov_f = 0.5 * (cov_f + cov_f.transpose(-1, -2))
    if q == 1:
        root = cov_f.clamp_min(0.0).sqrt()
    else:
        root = _chol(cov_f)
    samples = mean_f.unsqueeze(0) + (root.unsqueeze(0) @ eps_mc.unsqueeze(1).unsqueeze(-1)).squeeze(-1)  # (S,b,G,q)
    if kind == "raw":
        return samples
    if kind == "ucb":
        m = samples.mean(dim=0)
        return m + beta_prime * (samples - m).abs()
    if kind == "ei":
        obj = samples - best_f
        return softplus(torch.stack([obj, torch.zeros_like(obj)], dim=-1), -1, temp)
    raise ValueError(kind)


def construct_conformal_bands(gp: GPState, Xq, alpha, temp, grid_res, max_grid_refinements, eps_grid, z,
                              ratio=None, randomized=False, u_rand=None, lazy_prior_grad=False):
    """helpers.py:94-202 (t=1).  Returns target_grid (b,G,q), weights (b,G), conf_pred_mask (b,G,q),
    refit (dense caches of the kept pass), ood_mask (b,G,q)."""
    assert Xq.ndim >= 3 and grid_res >= 2
    b, q, _ = Xq.shape
    with torch.no_grad():                                            # helpers.py:118-119
        mean, cov = posterior(gp, Xq, observation_noise=True)
    if lazy_prior_grad and Xq.requires_grad:
        # GRADIENT SEMANTICS ONLY (values unchanged).  helpers.py:118-119 builds y_post under torch.no_grad(), but GPyTorch
        # returns the predictive covariance LAZILY: the prior block k(X, X) is a lazily evaluated kernel that is only
        # evaluated when the grid is sampled, after the no_grad block has exited, and so tracks d/dX; the correction term
        # k(X, X_train) A^-1 k(X_train, X) was solved for eagerly and is a constant (GPyTorch's LazyEvaluatedKernelTensor;
        # oracle/_shim restates exactly this).  For q = 1 the prior block is the constant outputscale and nothing leaks; for q >= 2 the grid
        # depends on X through k(x_j, x_k).  The product (and the default of this oracle) treats the grid as the constant
        # the no_grad block intends; this switch reproduces the reference's autograd to 1e-6 (tests/test_oracle_ref_path.py).
        Kxx = gp.k(Xq, Xq)
        cov = cov + (Kxx - Kxx.detach())
    if eps_grid.ndim == 2:
        eps_grid = eps_grid.unsqueeze(0).expand(b, -1, -1)           # collapse_batch_dims=True: shared over candidates
    best = (float("inf"), None, None, None, None, None)
    for refine_step in range(max_grid_refinements + 1):
        Lc = torch.linalg.cholesky(cov)                              # reads the lower triangle only
        target_grid = mean.unsqueeze(1) + (Lc.unsqueeze(1) @ eps_grid.unsqueeze(-1)).squeeze(-1)     # (b,G,q)
        # helpers.py:137  mvn.log_prob (joint over q)
        dev = target_grid - mean.unsqueeze(1)
        sol = torch.linalg.solve_triangular(Lc.unsqueeze(1), dev.unsqueeze(-1), upper=False).squeeze(-1)
        logdet = 2.0 * torch.diagonal(Lc, dim1=-2, dim2=-1).log().sum(-1)
        weights = -0.5 * (sol.pow(2).sum(-1) + logdet.unsqueeze(1) + q * LOG2PI)                     # (b,G)

        conf_pred_mask, refit, _, ood_mask = conformal_gp_regression(
            gp, Xq, target_grid, alpha, temp, z, ratio=ratio, randomized=randomized, u_rand=u_rand)

        num_accepted = (conf_pred_mask.detach() >= 0.5).float().sum(1)   # (b,q)  float32, as helpers.py:150-155
        min_acc, max_acc = num_accepted.min(), num_accepted.max()
        diff = num_accepted / grid_res - 0.5
        grid_score = diff.pow(2).mean()
        if grid_score < best[0]:
            best = (grid_score, target_grid, weights, conf_pred_mask, refit, ood_mask)
        covar_scale = 1.0 + diff.sign() * diff.abs().pow(2.0 / q)    # (b,q) float32 (helpers.py:166-167)
        if min_acc < 2 or max_acc > grid_res - 2:
            cov = covar_scale.unsqueeze(-1) * cov                    # helpers.py:170,178 (row scaling; promotes to fp64)
        else:
            break
    return best[1:]


def conformal_acquisition(gp: GPState, Xq, *, kind, alpha, temp, grid_res, max_grid_refinements=0,
                          eps_grid, z, eps_mc, ratio=None, randomized=False, u_rand=None,
                          optimistic=True, beta_prime=None, best_f=None, X_baseline=None, lazy_prior_grad=False):
    """ConformalAcquisition.forward (acquisitions.py:53-76).  Returns dict with acq (b,) and the bands.
    lazy_prior_grad: see construct_conformal_bands (gradient semantics of the reference's lazily evaluated grid covariance)."""
    b, q, _ = Xq.shape
    target_grid, logp, mask, refit, ood = construct_conformal_bands(
        gp, Xq, alpha, temp, grid_res, max_grid_refinements, eps_grid, z, ratio, randomized, u_rand,
        lazy_prior_grad=lazy_prior_grad)
    if optimistic:
        with torch.no_grad():                                        # acquisitions.py:58-59
            mean, _ = posterior(gp, Xq, observation_noise=False)
        opt_mask = (target_grid >= mean.unsqueeze(1)).to(target_grid)
    else:
        opt_mask = torch.ones_like(target_grid)
    if kind in ("nei", "joint"):
        values = nei_values(gp, refit, Xq, X_baseline, eps_mc, temp, raw=kind == "joint")
    else:
        values = nonconformal_values(gp, refit, q, eps_mc, kind, beta_prime=beta_prime, best_f=best_f, temp=temp)
    res = conformal_integration(values, mask.unsqueeze(-1), logp[..., None, None], alpha,
                                opt_mask.unsqueeze(-1), ood.unsqueeze(-1))          # (S,b,q)
    acq = res.max(dim=-1).values.mean(0)
    return dict(acq=acq, target_grid=target_grid, logp=logp, mask=mask, ood=ood, values=values)
