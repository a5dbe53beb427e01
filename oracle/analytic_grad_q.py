"""Hand-derived reverse pass of the closed-form path for JOINT batches q >= 1 (cUCB / cEI), CPU torch fp64, no autograd.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Generalises oracle/analytic_grad.py to the rank-q bordered update
(reference: conformalbo/helpers.py:218-333 with q > 1, acquisitions.py:53-76 `max` over q): this is
the algebra of `k_score_q_bwd` / `k_rows_bwd_q`, checked against torch.autograd through oracle/closed_form.py.  Constants
of the differentiation as in the reference: target grid, its log-density, the optimistic mask (helpers.py:118-119,
acquisitions.py:58-59), hard counts, the Bernoulli draws and the arg-max over q.

Stages
  R1 (per candidate batch; k_score_q_bwd)  ->  w_bar (q,N), mu_bar (q), ut_bar (q,Sp), drc_bar (q), S_bar (q,q)
     through per-train-point moments  m0 = sum_g Gam, mrr = sum_g Gam rp^2, mr = sum_g Gam rp, mrd_k = sum_g Gam rp dev_k
     (Gam_ig = sum_j Gam_ijg, rp = sqrt(iv_i) r_ig) and per-column totals of Gam_ijg {1, B_j^2, B_j, B_j dev_k}.
  R2 (per candidate point; GEMMs + k_rows_bwd_q)  u_bar_j = A^-1 (w_bar_j + z ut_bar_j) + mu_bar_j alpha - sum_k Ssym_jk w_k,
     x_bar_j = sum_i u_bar_ji dk(X_i,x_j)/dx_j + sum_{k != j} Ssym_jk dk(x_j,x_k)/dx_j + drc_bar_j d ratio/dx,  Ssym = S_bar + S_bar^T.
"""
from __future__ import annotations

import math

import torch

from .analytic_grad import _ratio_fwd_bwd
from .gp import GPState, SQRT5

LOG2PI = math.log(2.0 * math.pi)


def _matern_and_grad_factor(diff_scaled):
    """k(r)/outputscale and (dk/dr)/r/outputscale for r = |diff_scaled|."""
    r = diff_scaled.pow(2).sum(-1).clamp_min(1e-30).sqrt()
    e = torch.exp(-SQRT5 * r)
    return (1 + SQRT5 * r + 5.0 / 3.0 * r * r) * e, -(5.0 / 3.0) * (1 + SQRT5 * r) * e


def conformal_acquisition_grad_q(gp: GPState, Xq, *, kind, alpha, temp, grid_res, eps_grid, z, eps_mc, ratio=None,
                                 randomized=False, u_rand=None, optimistic=True, beta_prime=None, best_f=None,
                                 cscale=None):
    b, q, d = Xq.shape
    N, G, noise = gp.N, grid_res, gp.noise
    Sp = z.shape[-1]
    dt = Xq.dtype
    assert temp > 0, "soft path (the hard path has no gradient through the masks; see analytic_grad.py for q = 1)"
    eye = torch.eye(q, dtype=dt)
    ls = gp.lengthscale
    # ---- forward ----------------------------------------------------------------------------------------------------------
    diff = (Xq[:, :, None, :] - gp.X[None, None, :, :]) / ls               # (b,q,N,d)
    kU, gU = _matern_and_grad_factor(diff)
    U = gp.outputscale * kU                                                # (b,q,N)
    V = torch.linalg.solve_triangular(gp.L, U.transpose(-1, -2), upper=False)
    W = torch.linalg.solve_triangular(gp.L.T, V, upper=True).transpose(-1, -2)     # (b,q,N)
    mu = gp.mean + U @ gp.alpha
    diffq = (Xq[:, :, None, :] - Xq[:, None, :, :]) / ls                   # (b,q,q,d)
    kQ, gQ = _matern_and_grad_factor(diffq)
    Kqq = gp.outputscale * kQ
    S = Kqq + noise * eye - U @ W.transpose(-1, -2)
    Sinv = torch.linalg.inv(S)
    T = torch.cholesky_solve(z[:N], gp.L)
    Zc = z[N:]                                                             # (q,Sp)
    R = Zc[None] - U @ T                                                   # (b,q,Sp)
    C = Sinv @ R
    kvar = noise / (Sp + 1e-6)
    zcc = (Zc[None] * C).sum(-1)
    var_c = noise + kvar * (Sp - noise * zcc)                              # (b,q)
    zdot = C @ z[:N].T                                                     # (b,q,N)
    a = noise * gp.alpha
    h0 = 1.0 + (Sp - noise * (z[:N] * T).sum(-1)) / (Sp + 1e-6)
    var_i = noise * (h0[None] + kvar * (W * zdot).sum(1))                  # (b,N)
    iv_i, iv_c = 0.5 / var_i, 0.5 / var_c
    lv_i, lv_c = -0.5 * var_i.log() - 0.5 * LOG2PI, -0.5 * var_c.log() - 0.5 * LOG2PI
    if ratio is None:
        dr, drc, ddrc_dx = torch.ones(N, dtype=dt), torch.ones(b, q, dtype=dt), torch.zeros(b, q, d, dtype=dt)
    else:
        dr = ratio(gp.X).clamp_min(1e-6)
        drc, ddrc_dx = _ratio_fwd_bwd(ratio, Xq.reshape(-1, d))
        drc, ddrc_dx = drc.reshape(b, q), ddrc_dx.reshape(b, q, d)
    norm = dr.sum() + drc
    omega = drc / norm
    eps = eps_grid.unsqueeze(0).expand(b, -1, -1) if eps_grid.ndim == 2 else eps_grid        # (b,G,q)
    cov = S if cscale is None else cscale.unsqueeze(-1) * S
    Lc = torch.linalg.cholesky(cov)
    dev = (Lc.unsqueeze(1) @ eps.unsqueeze(-1)).squeeze(-1)                # (b,G,q) constants
    sol = torch.linalg.solve_triangular(Lc.unsqueeze(1), dev.unsqueeze(-1), upper=False).squeeze(-1)
    logp = -0.5 * (sol.pow(2).sum(-1) + 2.0 * torch.diagonal(Lc, dim1=-2, dim2=-1).log().sum(-1, keepdim=True) + q * LOG2PI)
    Y = mu.unsqueeze(1) + dev
    opt = (Y >= mu.unsqueeze(1)).to(dt) if optimistic else torch.ones_like(Y)
    kdens = 1.0 / logp.exp().clamp_min(1e-6)                               # (b,G)
    Bg = dev @ Sinv.transpose(-1, -2)                                      # (b,G,q)
    thr = lv_c[:, None, :] - iv_c[:, None, :] * (noise * Bg) ** 2          # (b,G,q)
    r_ig = a[None, None, :] - noise * (Bg @ W)                             # (b,G,N)
    sc = lv_i[:, None, :] - iv_i[:, None, :] * r_ig ** 2
    dd = thr[:, :, None, :] - sc[:, :, :, None]                            # (b,G,N,q)
    ddc = thr[:, :, :, None] - thr[:, :, None, :]                          # (b,G,j,k) = thr_j - thr_k
    n_g = (dd >= 0).sum(2).to(dt) + (ddc >= 0).sum(3).to(dt)               # (b,G,q)
    rho = torch.sigmoid(dd / temp)
    rhoc = torch.sigmoid(ddc / temp)
    srj = rho.sum(2) + rhoc.sum(3)
    sw = (rho * dr[None, None, :, None]).sum(2)
    scale = n_g / srj
    Mraw = scale * (sw + 0.5 * drc[:, None, :]) / norm[:, None, :]
    M = Mraw
    keep_live = torch.zeros(b, G, q, dtype=dt)
    if randomized:
        Dr = 0.5 * scale * omega[:, None, :]
        D = torch.full_like(M, 1e-6)
        D[..., q - 1] = Dr[..., q - 1].clamp_min(1e-6)                     # helpers.py:307: LAST row only
        pr = ((M - alpha) / D).clamp(0.0, 1.0)
        keep = 1.0 - (u_rand < pr).to(dt)
        M = M - keep * D
        keep_live[..., q - 1] = keep[..., q - 1] * (Dr[..., q - 1] > 1e-6).to(dt)
    mk = torch.sigmoid((M - alpha) / temp)
    ood = torch.sigmoid((omega - alpha) / temp)                            # (b,q)
    shr = eye - noise * Sinv
    mean_f = mu.unsqueeze(1) + dev @ shr.transpose(-1, -2)                 # (b,G,q)
    Cf = noise * shr
    Cf = 0.5 * (Cf + Cf.transpose(-1, -2))
    Lf = Cf.clamp_min(0).sqrt() if q == 1 else torch.linalg.cholesky(Cf)
    Sn = eps_mc.shape[0]
    f = mean_f.unsqueeze(0) + torch.einsum("bij,sgj->sbgi", Lf, eps_mc)    # (S,b,G,q)
    if kind == "ucb":
        m = f.mean(0)
        sgn = torch.sign(f - m)
        val = m + beta_prime * (f - m).abs()
    else:
        xx = (f - best_f) / temp
        val = temp * (torch.clamp_min(xx, 0) + torch.log1p(torch.exp(-xx.abs())))
        dval = torch.sigmoid(xx)
    cw, nc = mk * opt * kdens.unsqueeze(-1), (1 - mk) * opt                # (b,G,q)
    scw, snc = cw.sum(1), nc.sum(1)                                        # (b,q)
    dcw, dnc = scw.clamp_min(1e-6), snc.clamp_min(1e-6)
    A_s = (cw.unsqueeze(0) * val).sum(2)                                   # (S,b,q)
    B_s = (nc.unsqueeze(0) * val).sum(2)
    inner = (1 - alpha) * A_s / dcw + alpha * B_s / dnc                    # (S,b,q)
    res = (1 - ood) * inner
    best = res.max(-1)
    acq = best.values.mean(0)

    # ---- reverse R1 ---------------------------------------------------------------------------------------------------------
    kap = torch.zeros_like(res)
    kap.scatter_(-1, best.indices.unsqueeze(-1), 1.0 / Sn)                 # (S,b,q): d acq / d res
    one_ood = (1 - ood)
    val_bar = kap[:, :, None, :] * one_ood[None, :, None, :] * ((1 - alpha) * cw / dcw[:, None, :] + alpha * nc / dnc[:, None, :])[None]
    cw_bar = (kap[:, :, None, :] * one_ood[None, :, None, :] * (1 - alpha)
              * (val / dcw[None, :, None, :] - ((scw > 1e-6).to(dt) * 1.0)[None, :, None, :] * (A_s / dcw ** 2)[:, :, None, :])).sum(0)
    nc_bar = (kap[:, :, None, :] * one_ood[None, :, None, :] * alpha
              * (val / dnc[None, :, None, :] - ((snc > 1e-6).to(dt) * 1.0)[None, :, None, :] * (B_s / dnc ** 2)[:, :, None, :])).sum(0)
    ood_bar = -(kap * inner).sum(0)                                        # (b,q)
    mk_bar = (cw_bar * kdens.unsqueeze(-1) - nc_bar) * opt
    if kind == "ucb":
        m_bar = (val_bar * (1 - beta_prime * sgn)).sum(0)
        f_bar = val_bar * beta_prime * sgn + m_bar.unsqueeze(0) / Sn
    else:
        f_bar = val_bar * dval
    meanf_bar = f_bar.sum(0)                                               # (b,G,q)
    Lf_bar = torch.tril(torch.einsum("sbgj,sgk->bjk", f_bar, eps_mc))      # (b,q,q)
    mu_bar = meanf_bar.sum(1)                                              # (b,q)
    shr_bar = torch.einsum("bgj,bgk->bjk", meanf_bar, dev)
    dev_bar = meanf_bar @ shr                                              # dev_bar_gk = sum_j meanf_bar_gj shr_jk
    # Cholesky reverse: Cf_bar = Lf^-T Phi(Lf^T Lf_bar) Lf^-1, symmetrised
    if q == 1:
        Cf_bar = torch.where(Cf > 0, Lf_bar / (2 * Lf.clamp_min(1e-300)), torch.zeros_like(Cf))
    else:
        P_ = Lf.transpose(-1, -2) @ Lf_bar
        Phi = torch.tril(P_) - 0.5 * torch.diag_embed(torch.diagonal(P_, dim1=-2, dim2=-1))
        Li = torch.linalg.inv(Lf)
        Cf_bar = Li.transpose(-1, -2) @ Phi @ Li
        Cf_bar = 0.5 * (Cf_bar + Cf_bar.transpose(-1, -2))
    shr_bar = shr_bar + noise * Cf_bar
    Sinv_bar = -noise * shr_bar
    # masks -> P, Q
    M_bar = mk_bar * mk * (1 - mk) / temp
    omega_bar = ood_bar * ood * (1 - ood) / temp
    scale_bar = M_bar * (sw + 0.5 * drc[:, None, :]) / norm[:, None, :]
    D_bar = -keep_live * M_bar
    scale_bar = scale_bar + D_bar * 0.5 * omega[:, None, :]
    omega_bar = omega_bar + (D_bar * 0.5 * scale).sum(1)
    P = M_bar * scale / norm[:, None, :]                                   # (b,G,q)
    Q = -scale_bar * scale / srj
    drc_bar = (0.5 * P).sum(1)
    norm_bar = -(M_bar * Mraw).sum(1) / norm
    drc_bar = drc_bar + omega_bar / norm
    norm_bar = norm_bar - omega_bar * omega / norm
    drc_bar = drc_bar + norm_bar
    Gam = (P[:, :, None, :] * dr[None, None, :, None] + Q[:, :, None, :]) * rho * (1 - rho) / temp      # (b,G,N,q)
    Gamc = Q[:, :, :, None] * rhoc * (1 - rhoc) / temp                     # (b,G,j,k)
    thr_bar = Gam.sum(2) + Gamc.sum(3) - Gamc.sum(2)                       # (b,G,q)
    Gi = Gam.sum(3)                                                        # (b,G,N) = -sc_bar
    siv = iv_i.sqrt()
    rp = r_ig * siv[:, None, :]
    m0 = Gi.sum(1)                                                         # (b,N)
    mrr = (Gi * rp * rp).sum(1)
    mr = (Gi * rp).sum(1)
    mrd = torch.einsum("bgi,bgk->bik", Gi * rp, dev)                       # (b,N,q)
    lv_i_bar = -m0
    iv_i_bar = mrr / iv_i
    var_i_bar = -2 * iv_i ** 2 * iv_i_bar - iv_i * lv_i_bar
    mrB = mrd @ Sinv.transpose(-1, -2)                                     # (b,N,q): sum_g Gam rp B_j
    w_bar = (-2 * noise * siv[:, :, None] * mrB).transpose(-1, -2)         # (b,q,N)
    w_bar = w_bar + (var_i_bar * noise * kvar)[:, None, :] * zdot
    zdot_bar = (var_i_bar * noise * kvar)[:, None, :] * W                  # (b,q,N)
    C_bar = zdot_bar @ z[:N]                                               # (b,q,Sp)
    # B_bar: train part through moments, candidate part through thr
    Bsum_bar = -2 * noise * torch.einsum("bi,bji,bi->bj", siv, W, mr)      # sum_g B_bar_gj (train)
    Sinv_bar = Sinv_bar - 2 * noise * torch.einsum("bi,bji,bik->bjk", siv, W, mrd)
    t_c = -2 * iv_c[:, None, :] * noise ** 2 * Bg * thr_bar                # B_bar_gj (threshold part)
    Bsum_bar = Bsum_bar + t_c.sum(1)
    Sinv_bar = Sinv_bar + torch.einsum("bgj,bgk->bjk", t_c, dev)
    lv_c_bar = thr_bar.sum(1)
    iv_c_bar = -(noise ** 2) * (thr_bar * Bg ** 2).sum(1)
    var_c_bar = -2 * iv_c ** 2 * iv_c_bar - iv_c * lv_c_bar
    C_bar = C_bar + (-var_c_bar * kvar * noise)[:, :, None] * Zc[None]
    # dev = Y - mu:  mu_bar_k -= sum_g dev_bar_gk,  dev_bar_gk = (mean_f part) + sum_j Sinv_jk B_bar_gj
    mu_bar = mu_bar - dev_bar.sum(1) - torch.einsum("bjk,bj->bk", Sinv, Bsum_bar)
    # C = Sinv R
    Sinv_bar = Sinv_bar + C_bar @ R.transpose(-1, -2)
    ut_bar = -(Sinv.transpose(-1, -2) @ C_bar)                             # (b,q,Sp)
    S_bar = -Sinv.transpose(-1, -2) @ Sinv_bar @ Sinv.transpose(-1, -2)
    Ssym = S_bar + S_bar.transpose(-1, -2)

    # ---- reverse R2 ---------------------------------------------------------------------------------------------------------
    rhs = w_bar + ut_bar @ z[:N].T                                         # fold the probe term: T = A^-1 z
    Aw = torch.cholesky_solve(rhs.reshape(b * q, N).T.contiguous(), gp.L).T.reshape(b, q, N)
    u_bar = Aw + mu_bar[:, :, None] * gp.alpha[None, None, :] - Ssym @ W
    grad = ((u_bar * gp.outputscale * gU)[..., None] * diff / ls).sum(2)   # (b,q,d)
    offd = Ssym * (1 - eye)                                                # the diagonal of K_qq is a constant
    grad = grad + ((offd * gp.outputscale * gQ)[..., None] * diffq / ls).sum(2)
    grad = grad + drc_bar[:, :, None] * ddrc_dx
    return dict(acq=acq, grad=grad, w_bar=w_bar, mu_bar=mu_bar, ut_bar=ut_bar, drc_bar=drc_bar, S_bar=S_bar)
