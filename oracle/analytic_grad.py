"""Hand-derived reverse pass of the closed-form path (q = 1, cUCB / cEI), CPU torch fp64, no autograd.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is the algebra the CUDA backward kernels
(`score_bwd.cuh`, `rows_bwd.cuh`) implement, written with dense (b, G, N) tensors so that it can be checked
against `torch.autograd` through `oracle/closed_form.py` / `oracle/dense.py` (tests/test_oracle_consistency.py) and
then serve as a stage-by-stage checker of the kernels.  What the reference differentiates: `optimize_acqf_sgld`
calls `acq_function(X).sum().backward()` (conformalbo/optim/optimize.py:224-235); the target grid, its log-density
and the optimistic mask are built under `torch.no_grad()` (helpers.py:118-119, acquisitions.py:58-59) and are
constants of the differentiation.

Stages (names as in SURVEY.md App. A):
  forward   u, w = A^-1 u, mu, s, ut = u^T T, drc -> per-train-point quadratic score difference
            Delta_ig = A_i b_g^2 + B_i b_g + C_i, b_g = (Y_g - mu)/s -> rho, mass M_g, mask, weights -> acq
  reverse   d acq/d(mask, values) -> per-grid P_g, Q_g (d acq/d rho_ig = P_g dr_i + Q_g) ->
            E_k,i = sum_g Gamma_ig b_g^k (k = 0,1,2), Gamma = (P dr_i + Q) rho (1 - rho) / temp ->
            w_bar, s_bar, mu_bar, ut_bar, drc_bar -> u_bar = A^-1 w_bar + mu_bar alpha - 2 s_bar w + T ut_bar ->
            x_bar = sum_i u_bar_i dk(X_i, x)/dx + drc_bar d drc/dx
"""
from __future__ import annotations

import math

import torch

from .gp import GPState, SQRT5

LOG2PI = math.log(2.0 * math.pi)


def _ratio_fwd_bwd(ratio, x):
    """dr(x) = clamp_min(min(p, 1-1e-6)/max(1-p, 1e-6), 1e-6) and d dr/d x for the d->H->H->1 ReLU MLP."""
    a1 = x @ ratio.W1.T + ratio.b1
    h1 = a1.clamp_min(0)
    a2 = h1 @ ratio.W2.T + ratio.b2
    h2 = a2.clamp_min(0)
    logit = (h2 @ ratio.W3.T + ratio.b3).squeeze(-1)
    p = torch.sigmoid(logit)
    num, den = p.clamp_max(1 - 1e-6), (1 - p).clamp_min(1e-6)
    w = num / den
    dw_dp = (p <= 1 - 1e-6).to(p) / den + ((1 - p) >= 1e-6).to(p) * num / den.pow(2)
    dw_dp = dw_dp * (w >= 1e-6).to(p)
    dlogit = dw_dp * p * (1 - p)                                           # (b,)
    g2 = (dlogit.unsqueeze(-1) * ratio.W3) * (a2 > 0).to(p)                # (b,H)
    g1 = (g2 @ ratio.W2) * (a1 > 0).to(p)
    return w.clamp_min(1e-6), g1 @ ratio.W1                                # (b,), (b,d)


def conformal_acquisition_grad(gp: GPState, Xq, *, kind, alpha, temp, grid_res, eps_grid, z, eps_mc, ratio=None,
                               randomized=False, u_rand=None, optimistic=True, beta_prime=None, best_f=None,
                               cscale=None):
    """Returns dict(acq (b,), grad (b,1,d), and the intermediate adjoints w_bar, s_bar, mu_bar, ut_bar, drc_bar)."""
    b, q, d = Xq.shape
    assert q == 1
    N, G, noise = gp.N, grid_res, gp.noise
    Sp = z.shape[-1]
    x = Xq[:, 0, :]
    dt = x.dtype
    # ---- forward ------------------------------------------------------------------------------
    diff = (x[:, None, :] - gp.X[None, :, :]) / gp.lengthscale            # (b,N,d)
    r = diff.pow(2).sum(-1).clamp_min(1e-30).sqrt()
    e = torch.exp(-SQRT5 * r)
    U = gp.outputscale * (1 + SQRT5 * r + 5.0 / 3.0 * r * r) * e          # (b,N)
    V = torch.linalg.solve_triangular(gp.L, U.T, upper=False)             # (N,b)
    W = torch.linalg.solve_triangular(gp.L.T, V, upper=True).T            # (b,N)
    mu = gp.mean + U @ gp.alpha
    s = gp.outputscale + noise - (V * V).sum(0)
    T = torch.cholesky_solve(z[:N], gp.L)                                 # (N,Sp)
    ut = U @ T                                                            # (b,Sp)
    zc = z[N]                                                             # (Sp,)
    if ratio is None:
        dr = torch.ones(N, dtype=dt)
        drc = torch.ones(b, dtype=dt)
        ddrc_dx = torch.zeros(b, d, dtype=dt)
    else:
        dr = ratio(gp.X).clamp_min(1e-6)
        drc, ddrc_dx = _ratio_fwd_bwd(ratio, x)
    sum_dr = dr.sum()
    if eps_grid.ndim == 2:
        eps = eps_grid[:, 0].unsqueeze(0).expand(b, -1)
    else:
        eps = eps_grid[:, :, 0]
    cs = torch.ones(b, dtype=dt) if cscale is None else cscale
    gsd = (cs * s).sqrt()
    inv_s = 1.0 / s
    bg = gsd[:, None] * eps * inv_s[:, None]                              # (b,G)
    Y = mu[:, None] + gsd[:, None] * eps
    logp = -0.5 * (eps * eps + (cs * s).log()[:, None] + LOG2PI)
    opt = (Y >= mu[:, None]).to(dt) if optimistic else torch.ones_like(Y)
    kappa_g = opt / logp.exp().clamp_min(1e-6)

    kvar = noise / (Sp + 1e-6)
    c = (zc[None, :] - ut) * inv_s[:, None]                               # (b,Sp)
    zcc = c @ zc
    var_c = noise + kvar * (Sp - noise * zcc)                             # (b,)
    zdot = c @ z[:N].T                                                    # (b,N)
    a = noise * gp.alpha
    h0 = 1.0 + (Sp - noise * (z[:N] * T).sum(-1)) / (Sp + 1e-6)
    var_i = noise * (h0[None, :] + kvar * W * zdot)                       # (b,N)
    iv_i, iv_c = 0.5 / var_i, 0.5 / var_c
    lv_i, lv_c = -0.5 * var_i.log() - 0.5 * LOG2PI, -0.5 * var_c.log() - 0.5 * LOG2PI
    wi = noise * W
    A_i = iv_i * wi * wi - (iv_c * noise * noise)[:, None]
    B_i = -2.0 * iv_i * a[None, :] * wi
    C_i = iv_i * a[None, :] ** 2 + lv_c[:, None] - lv_i
    Delta = (A_i[:, None, :] * bg[:, :, None] + B_i[:, None, :]) * bg[:, :, None] + C_i[:, None, :]   # (b,G,N)
    n_g = (Delta >= 0).sum(-1).to(dt) + 1.0
    norm = sum_dr + drc
    omega = drc / norm
    soft = temp > 0
    if soft:
        rho = torch.sigmoid(Delta / temp)
        sr = rho.sum(-1)
        sw = (rho * dr).sum(-1)
        scale = n_g / (sr + 0.5)
        rho_self = 0.5 * scale
        Mraw = scale * (sw + 0.5 * drc[:, None]) / norm[:, None]
    else:
        sw = ((Delta >= 0).to(dt) * dr).sum(-1)
        rho_self = torch.ones(b, G, dtype=dt)
        Mraw = (sw + drc[:, None]) / norm[:, None]
    M = Mraw
    keep = torch.zeros(b, G, dtype=dt)
    if randomized:
        Dg = (rho_self * omega[:, None]).clamp_min(1e-6)
        pr = ((M - alpha) / Dg).clamp(0.0, 1.0)
        keep = 1.0 - (u_rand.reshape(b, G) < pr).to(dt)
        M = M - keep * Dg
    mk = torch.sigmoid((M - alpha) / temp) if soft else (M > alpha).to(dt)
    ood = torch.sigmoid((omega - alpha) / temp) if soft else (omega >= alpha).to(dt)
    cw, nc = mk * kappa_g, (1 - mk) * opt
    scw, snc = cw.sum(-1), nc.sum(-1)
    shrink = 1.0 - noise * inv_s
    mf = mu[:, None] + (Y - mu[:, None]) * shrink[:, None]
    sd_f = (noise * shrink).clamp_min(0).sqrt()
    em = eps_mc.reshape(-1, G)                                            # (S,G)
    S = em.shape[0]
    f = mf[None] + sd_f[None, :, None] * em[:, None, :]                   # (S,b,G)
    if kind == "ucb":
        ebar = em.mean(0)
        coef = ebar[None, :] + beta_prime * (em - ebar[None, :]).abs()    # (S,G): val = mf + sd_f * coef
        val = mf[None] + sd_f[None, :, None] * coef[:, None, :]
        dval_dmf = torch.ones_like(val)
        dval_dsd = coef[:, None, :].expand(S, b, G)
    else:
        xx = f - best_f
        if soft:
            val = temp * (torch.clamp_min(xx / temp, 0) + torch.log1p(torch.exp(-(xx / temp).abs())))
            dval_dmf = torch.sigmoid(xx / temp)
        else:
            val = xx.clamp_min(0)
            dval_dmf = (xx > 0).to(dt)
        dval_dsd = dval_dmf * em[:, None, :]
    Asum = (cw[None] * val).sum(-1).sum(0)                                # (b,)  sum over s of A_s
    Bsum = (nc[None] * val).sum(-1).sum(0)
    dcw, dnc = scw.clamp_min(1e-6), snc.clamp_min(1e-6)
    inner = ((1 - alpha) * Asum / dcw + alpha * Bsum / dnc) / S
    acq = (1 - ood) * inner

    # ---- reverse ------------------------------------------------------------------------------
    ka = (1 - ood) * (1 - alpha) / dcw / S
    kb = (1 - ood) * alpha / dnc / S
    val_bar = ka[None, :, None] * cw[None] + kb[None, :, None] * nc[None]          # (S,b,G)
    vsum = val.sum(0)                                                               # (b,G)
    scw_bar = torch.where(scw > 1e-6, -ka * Asum / dcw, torch.zeros_like(scw))
    snc_bar = torch.where(snc > 1e-6, -kb * Bsum / dnc, torch.zeros_like(snc))
    cw_bar = ka[:, None] * vsum + scw_bar[:, None]
    nc_bar = kb[:, None] * vsum + snc_bar[:, None]
    mk_bar = cw_bar * kappa_g - nc_bar * opt
    ood_bar = -inner
    mf_bar = (val_bar * dval_dmf).sum(0)                                            # (b,G)
    sd_bar = (val_bar * dval_dsd).sum((0, 2))                                       # (b,)
    drc_bar = torch.zeros(b, dtype=dt)
    w_bar = torch.zeros(b, N, dtype=dt)
    s_bar = torch.zeros(b, dtype=dt)
    mu_bar = torch.zeros(b, dtype=dt)
    ut_bar = torch.zeros(b, Sp, dtype=dt)
    Sb0 = torch.zeros(b, dtype=dt)      # sum_g b_bar_g
    Sb1 = torch.zeros(b, dtype=dt)      # sum_g b_bar_g b_g
    if soft:
        M_bar = mk_bar * mk * (1 - mk) / temp
        omega_bar = ood_bar * ood * (1 - ood) / temp
        scale_bar = M_bar * (sw + 0.5 * drc[:, None]) / norm[:, None]
        if randomized:
            D_bar = -keep * M_bar
            live = (rho_self * omega[:, None] > 1e-6).to(dt)
            scale_bar = scale_bar + D_bar * live * 0.5 * omega[:, None]
            omega_bar = omega_bar + (D_bar * live * rho_self).sum(-1)
        P = M_bar * scale / norm[:, None]                                           # d/d sw_g
        drc_bar = drc_bar + (0.5 * P).sum(-1)
        norm_bar = -(M_bar * Mraw).sum(-1) / norm
        Q = -scale_bar * scale / (sr + 0.5)                                         # d/d sr_g
        drc_bar = drc_bar + omega_bar / norm
        norm_bar = norm_bar - omega_bar * omega / norm
        drc_bar = drc_bar + norm_bar
        Gam = (P[:, :, None] * dr[None, None, :] + Q[:, :, None]) * rho * (1 - rho) / temp     # (b,G,N)
        E0 = Gam.sum(1)
        E1 = (Gam * bg[:, :, None]).sum(1)
        E2 = (Gam * (bg * bg)[:, :, None]).sum(1)                                    # (b,N)
        Sb0 = (2 * A_i * E1 + B_i * E0).sum(-1)
        Sb1 = (2 * A_i * E2 + B_i * E1).sum(-1)
        iv_i_bar = E2 * wi * wi - 2 * E1 * a[None, :] * wi + E0 * a[None, :] ** 2
        w_bar = 2 * iv_i * noise * (E2 * wi - E1 * a[None, :])
        iv_c_bar = -noise * noise * E2.sum(-1)
        lv_c_bar = E0.sum(-1)
        var_i_bar = -2 * iv_i * iv_i * iv_i_bar + iv_i * E0
        var_c_bar = -2 * iv_c * iv_c * iv_c_bar - iv_c * lv_c_bar
        w_bar = w_bar + var_i_bar * noise * kvar * zdot
        zdot_bar = var_i_bar * noise * kvar * W                                     # (b,N)
        c_bar = zdot_bar @ z[:N] + (-var_c_bar * kvar * noise)[:, None] * zc[None, :]          # (b,Sp)
        ut_bar = -c_bar * inv_s[:, None]
        s_bar = s_bar - (c_bar * c).sum(-1) * inv_s
    # refit mean / sd and the grid variable b_g = (Y_g - mu)/s
    P_f = s - noise
    Sb0 = Sb0 + P_f * mf_bar.sum(-1)
    Sb1 = Sb1 + P_f * (mf_bar * bg).sum(-1)
    mu_bar = mu_bar + mf_bar.sum(-1) - Sb0 * inv_s
    s_bar = s_bar + (mf_bar * bg).sum(-1) - Sb1 * inv_s
    s_bar = s_bar + torch.where(sd_f > 0, sd_bar * noise * noise * inv_s * inv_s / (2 * sd_f.clamp_min(1e-300)),
                                torch.zeros_like(sd_f))
    # kernel rows: u_bar = A^-1 w_bar + mu_bar alpha - 2 s_bar w + T ut_bar
    Aw = torch.cholesky_solve(w_bar.T.contiguous(), gp.L).T                          # (b,N)
    u_bar = Aw + mu_bar[:, None] * gp.alpha[None, :] - 2 * s_bar[:, None] * W + ut_bar @ T.T
    dk_dr_over_r = -(5.0 / 3.0) * gp.outputscale * (1 + SQRT5 * r) * e                # dk/dr / r
    grad = ((u_bar * dk_dr_over_r)[:, :, None] * diff / gp.lengthscale).sum(1)        # (b,d)
    grad = grad + drc_bar[:, None] * ddrc_dx
    return dict(acq=acq, grad=grad.unsqueeze(1), w_bar=w_bar, s_bar=s_bar, mu_bar=mu_bar, ut_bar=ut_bar,
                drc_bar=drc_bar, u_bar=u_bar)
