"""Restatements of the reference's pure-torch helpers (test infrastructure only).

These four are the only pieces of the path whose reference implementation can be executed in
this image (no botorch/gpytorch needed); `oracle/make_golden.py` runs the originals from
/root/reference and `tests/test_oracle_golden.py` pins the functions below against those outputs.
"""
from __future__ import annotations

import torch


def softplus(x: torch.Tensor, dim: int, temp: float, keepdim: bool = False) -> torch.Tensor:
    """temp * logsumexp(x / temp)  -- conformalbo/acquisition/monte_carlo.py:116-117."""
    return temp * torch.logsumexp(x / temp, dim, keepdim=keepdim)


def conf_mask_to_bounds(target_grid: torch.Tensor, conf_pred_mask: torch.Tensor):
    """Band endpoints from a (soft) conformal mask -- conformalbo/helpers.py:15-91.

    target_grid, conf_pred_mask: (b, G, 1, t).  Returns lb, ub: (b, 1, t).
    Per (candidate, output): sort the grid ascending, take the first / last sorted position whose
    mask is >= 0.5 and move the endpoint to the linear-interpolated 0.5 crossing towards the
    neighbouring rejected grid point (unless the accepted point is the grid edge).  No accepted
    point -> (+inf, +inf)  (helpers.py:48-51: `ones/0.`).
    """
    assert target_grid.shape == conf_pred_mask.shape and target_grid.shape[-2] == 1
    b, G, _, t = target_grid.shape
    grid = target_grid[:, :, 0, :].transpose(1, 2).reshape(b * t, G)   # rows = (cand, output)
    mask = conf_pred_mask[:, :, 0, :].transpose(1, 2).reshape(b * t, G)
    grid, order = torch.sort(grid, dim=-1)
    mask = torch.gather(mask, 1, order)
    acc = mask >= 0.5
    any_acc = acc.any(-1)
    pos = torch.arange(G).expand_as(acc)
    first = torch.where(acc, pos, torch.full_like(pos, G)).min(-1).values.clamp_max(G - 1)
    last = torch.where(acc, pos, torch.full_like(pos, -1)).max(-1).values.clamp_min(0)
    rows = torch.arange(b * t)

    def at(x, idx):
        return x[rows, idx.clamp(0, G - 1)]

    g_lo, m_lo = at(grid, first), at(mask, first)
    g_lo_prev, m_lo_prev = at(grid, first - 1), at(mask, first - 1)
    lb_interp = g_lo_prev + (0.5 - m_lo_prev) * (g_lo - g_lo_prev) / (m_lo - m_lo_prev)
    lb = torch.where(first != 0, lb_interp, g_lo)

    g_hi, m_hi = at(grid, last), at(mask, last)
    g_hi_next, m_hi_next = at(grid, last + 1), at(mask, last + 1)
    ub_interp = g_hi + (m_hi - 0.5) * (g_hi_next - g_hi) / (m_hi - m_hi_next)
    ub = torch.where(last != G - 1, ub_interp, g_hi)

    inf = torch.full_like(lb, float("inf"))
    lb = torch.where(any_acc, lb, inf).reshape(b, t).unsqueeze(1)
    ub = torch.where(any_acc, ub, inf).reshape(b, t).unsqueeze(1)
    return lb, ub


def conformal_integration(values, conf_pred_mask, grid_logp, alpha, opt_mask, ood_mask):
    """Importance-weighted reduction over the outcome grid -- conformalbo/acquisitions.py:225-248.

    values (S, b, G, q); masks (b, G, q, t); grid_logp (b, G, 1, 1)  ->  (S, b, q).
    Trailing dims of the masks beyond `values.ndim - 1` are collapsed (product for masks, sum for
    the log-density), then over the grid axis:
      nonconf_w = (1-mask)*opt / max(sum, 1e-6);  conf_w = mask*opt/max(exp(logp),1e-6) / max(sum, 1e-6)
      out = sum_g (1-ood) * ((1-alpha) conf_w + alpha nonconf_w) * values
    """
    n_collapse = conf_pred_mask.ndim - (values.ndim - 1)
    for _ in range(n_collapse):
        opt_mask = opt_mask.prod(-1)
        conf_pred_mask = conf_pred_mask.prod(-1)
        ood_mask = ood_mask.prod(-1)
        grid_logp = grid_logp.sum(-1)
    gdim = -3 + n_collapse
    nonconf = (1.0 - conf_pred_mask) * opt_mask
    nonconf = nonconf / nonconf.sum(gdim, keepdim=True).clamp_min(1e-6)
    conf = (conf_pred_mask * opt_mask) / grid_logp.exp().clamp_min(1e-6)
    conf = conf / conf.sum(gdim, keepdim=True).clamp_min(1e-6)
    w = (1.0 - alpha) * conf + alpha * nonconf
    return ((1.0 - ood_mask) * w * values).sum(gdim)
