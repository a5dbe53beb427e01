"""CPU: both oracle restatements against goldens produced by EXECUTING the reference's own helpers.py / acquisitions.py
(oracle/make_golden_path.py: reference source unmodified over the gpytorch/botorch stand-ins of oracle/_shim).

Tolerances are the north star's: acquisition values, band endpoints, grid 1e-6 relative; soft masks 1e-6 absolute;
arg-max index identical."""
import math
import os

import pytest
import torch

from oracle import closed_form, dense, pure
from oracle import gp as ogp

GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-6


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def acq_cases(name="ref_path.pt"):
    return load(name)["acq"]


def oracle_inputs(c):
    p = c["problem"]
    gp = ogp.build_state(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"])
    Xq = c["Xq"] if c["X_pending"] is None else torch.cat([c["Xq"], c["X_pending"].expand(c["b"], -1, -1)], dim=1)
    ratio = ogp.RatioMLP(*c["ratio_weights"]) if c["ratio"] else None
    kw = dict(kind=c["kind"], alpha=c["alpha"], temp=c["temp"], grid_res=c["G"], max_grid_refinements=c["refine"],
              eps_grid=c["eps_grid"], z=c["z"], eps_mc=c["eps_mc"], ratio=ratio, randomized=c["randomized"],
              u_rand=c["u_rand"], optimistic=c["optimistic"], beta_prime=math.sqrt(c["beta"] * math.pi / 2),
              best_f=c["best_f"], X_baseline=c["X_baseline"])
    return gp, Xq, kw


def check_against_golden(c, got, mask_atol=1e-6):
    """got: dict with acq (b,), target_grid (b,G,q), logp (b,G), mask (b,G,q), ood (b,G,q) or (b,q)."""
    want = c["out"]
    b, G = c["b"], c["G"]
    assert torch.allclose(got["acq"], want["acq"], rtol=RTOL, atol=1e-9), (got["acq"] - want["acq"]).abs().max()
    assert int(got["acq"].argmax()) == want["argmax"]
    assert torch.allclose(got["target_grid"], want["target_grid"][..., 0], rtol=RTOL, atol=1e-9)
    assert torch.allclose(got["logp"], want["grid_logp"], rtol=RTOL, atol=1e-9)
    wm = want["conf_pred_mask"][..., 0]
    if c["temp"] > 0:
        assert torch.allclose(got["mask"], wm, rtol=0, atol=mask_atol), (got["mask"] - wm).abs().max()
    else:
        assert torch.equal(got["mask"], wm)
    ood = got["ood"] if got["ood"].ndim == 3 else got["ood"].unsqueeze(1).expand(b, G, -1)
    assert torch.allclose(ood, want["ood_mask"][..., 0], rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("idx", range(12))
def test_closed_form_against_reference_source_golden(idx):
    c = acq_cases()[idx]
    gp, Xq, kw = oracle_inputs(c)
    got = closed_form.conformal_acquisition(gp, Xq, **kw)
    check_against_golden(c, got)
    if "lb" in c["out"]:
        lb, ub = pure.conf_mask_to_bounds(got["target_grid"].unsqueeze(-1), got["mask"].unsqueeze(-1))
        fin = torch.isfinite(c["out"]["lb"])
        assert torch.equal(torch.isfinite(lb.reshape(-1)), fin)
        if c["temp"] > 1e-3:
            assert torch.allclose(lb.reshape(-1)[fin], c["out"]["lb"][fin], rtol=RTOL, atol=1e-7)
            assert torch.allclose(ub.reshape(-1)[fin], c["out"]["ub"][fin], rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("idx", range(12))
def test_dense_against_reference_source_golden(idx):
    c = acq_cases()[idx]
    gp, Xq, kw = oracle_inputs(c)
    got = dense.conformal_acquisition(gp, Xq, **kw)
    check_against_golden(c, got)


def test_golden_file_covers_the_named_configs():
    names = [c["name"] for c in acq_cases()]
    assert len(names) == 12
    for want in ("cfg1_ucb_q1", "cfg1_ucb_q2", "cfg4_pool", "small_nei", "small_q3", "small_pending"):
        assert want in names
    big = [c["name"] for c in acq_cases("ref_path_n1024.pt")]
    assert "cfg2_ucb" in big and "cfg2_ucb_q2" in big


@pytest.mark.parametrize("idx", range(3))
def test_closed_form_against_reference_source_golden_n1024(idx):
    """BASELINE configs[1] sizes (N=1024, d=8, G=64): 48 candidates scored by the reference's dense algorithm."""
    c = acq_cases("ref_path_n1024.pt")[idx]
    gp, Xq, kw = oracle_inputs(c)
    got = closed_form.conformal_acquisition(gp, Xq, **kw)
    check_against_golden(c, got)


def _multi_gps(c):
    p = c["problem"]
    t = c["t"]
    return [ogp.build_state(p["X"], p["Y"][:, k], float(p["mean"][k]), p["lengthscale"][k], float(p["outputscale"][k]),
                            float(p["noise"][k])) for k in range(t)]


@pytest.mark.parametrize("name", ["cov_t2", "cov_t2_rand"])
def test_multi_output_bands_against_reference_source_golden(name):
    """t = 2: oracle/multi.py with label_layout='reference' (the `.view` scramble of helpers.py:257-258)
    reproduces what helpers.assess_coverage itself computes; label_layout='intent' differs (that is the point of the flag)."""
    from oracle import multi
    c = [x for x in load("ref_path.pt")["coverage"] if x["name"] == name][0]
    gps = _multi_gps(c)
    ratio = ogp.RatioMLP(*c["ratio_weights"]) if c["ratio"] else None
    kw = dict(ratio=ratio, randomized=c["randomized"], u_rand=c["u_rand"])
    std_cov, conf_cov, grid, mask, lb, ub = multi.assess_coverage(
        gps, c["X_holdout"], c["Y_holdout"], c["alpha"], c["temp"], c["G"], c["refine"], c["eps_grid"], c["z"],
        label_layout="reference", **kw)
    want = c["out"]
    assert abs(std_cov - want["std_coverage"]) < 1e-6 and abs(conf_cov - want["conformal_coverage"]) < 1e-6
    assert torch.allclose(grid, want["target_grid"], rtol=RTOL, atol=1e-9)
    assert torch.allclose(mask, want["conf_pred_mask"], rtol=0, atol=1e-6), (mask - want["conf_pred_mask"]).abs().max()
    fin = torch.isfinite(want["lb"])
    assert torch.equal(torch.isfinite(lb), fin)
    assert torch.allclose(lb[fin], want["lb"][fin], rtol=RTOL, atol=1e-7) and torch.allclose(ub[fin], want["ub"][fin], rtol=RTOL, atol=1e-7)
    _, _, grid_i, mask_i, _, _ = multi.assess_coverage(
        gps, c["X_holdout"], c["Y_holdout"], c["alpha"], c["temp"], c["G"], c["refine"], c["eps_grid"], c["z"],
        label_layout="intent", **kw)
    assert (mask_i - mask).abs().max() > 1e-3          # the scramble is not a no-op


def test_multi_output_intent_layout_is_t_independent_single_output_problems():
    """label_layout='intent' without refinement == the single-output oracle per output (stacked); with refinement the
    outputs are tied through the joint accept statistics only."""
    from oracle import multi
    c = [x for x in load("ref_path.pt")["coverage"] if x["name"] == "cov_t2"][0]
    gps = _multi_gps(c)
    X = c["X_holdout"][:6, None]
    eps = c["eps_grid"].movedim(0, 1)[:6]                                   # (n, G, 1, t)
    grid, logp, mask, ood = multi.construct_conformal_bands(gps, X, c["alpha"], c["temp"], c["G"], 0, eps, c["z"],
                                                            label_layout="intent")
    lp = 0.0
    for k, gp in enumerate(gps):
        g1, l1, m1, _, o1 = dense.construct_conformal_bands(gp, X, c["alpha"], c["temp"], c["G"], 0, eps[..., k], c["z"])
        assert torch.allclose(grid[..., k], g1) and torch.allclose(mask[..., k], m1) and torch.allclose(ood[..., k], o1)
        lp = lp + l1
    assert torch.allclose(logp, lp)


@pytest.mark.parametrize("name", ["ehvi_q1", "ehvi_q2_rand"])
def test_cehvi_inner_expectation_against_reference_source_golden(name):
    """cEHVI (mobo_acquisitions.py:24-45) on a two-output model.  The golden holds the ARGUMENTS with which the reference's
    own `_conformal_fwd` calls `_conformal_integration` (oracle/make_golden_path.py::run_ehvi_case): `values` = the
    reference's `_nonconformal_fwd` -- `_compute_qehvi` under every refit model, the cEHVI inner expectation (b, G) -- the
    conformal masks, the joint grid log-density, the optimistic mask.  oracle/mobo.py reproduces values / log-density /
    optimistic mask, oracle/multi.py (label_layout='reference') the masks.  The integration that follows is unpinned: the
    reference's function raises for these shapes."""
    from oracle import mobo, multi
    c = [x for x in load("ref_path_ehvi.pt")["ehvi"] if x["name"] == name][0]
    gps = _multi_gps(c)
    ratio = ogp.RatioMLP(*c["ratio_weights"]) if c["ratio"] else None
    want = c["out"]
    b, q, G = c["b"], c["q"], c["G"]
    # A non-interleaved MultitaskMultivariateNormal consumes its (..., q, t) base samples as the FLAT task-major event
    # (oracle/_shim/gpytorch/distributions.py): entry (point j, task k) gets flat[k * q + j]
    remap = lambda e: e.reshape(*e.shape[:-2], e.shape[-1], e.shape[-2]).transpose(-1, -2)
    eps_grid = remap(c["eps_grid"][:, 0])                                   # (G, q, t)
    eps_mc = remap(c["eps_mc"][:, 0])                                       # (S, G, q, t)
    u_qm = c["u_rand"].permute(0, 1, 3, 2) if c["randomized"] else None     # (b, G, t, q) -> (b, G, q, m)
    got = mobo.cehvi(gps, c["Xq"], alpha=c["alpha"], temp=c["temp"], grid_res=G, eps_grid=eps_grid, z=c["z"], eps_mc=eps_mc,
                     cell_lower=c["cell_lower"], cell_upper=c["cell_upper"], ratio=ratio, randomized=c["randomized"],
                     u_rand=u_qm, optimistic=c["optimistic"], impl="dense")
    assert want["values"].abs().max() > 1e-2
    assert torch.allclose(got["values"], want["values"], rtol=RTOL, atol=1e-9), (got["values"] - want["values"]).abs().max()
    assert torch.allclose(got["logp"], want["grid_logp"].reshape(b, G), rtol=RTOL, atol=1e-9)
    closed = mobo.cehvi(gps, c["Xq"], alpha=c["alpha"], temp=c["temp"], grid_res=G, eps_grid=eps_grid, z=c["z"], eps_mc=eps_mc,
                        cell_lower=c["cell_lower"], cell_upper=c["cell_upper"], ratio=ratio, randomized=c["randomized"],
                        u_rand=u_qm, optimistic=c["optimistic"], impl="closed")
    assert torch.allclose(closed["values"], want["values"], rtol=RTOL, atol=1e-9)
    # the masks the reference hands to its integration are those of its (label-scrambling) multi-output band construction
    grid, logp, mask, ood = multi.construct_conformal_bands(gps, c["Xq"], c["alpha"], c["temp"], G, 0, eps_grid, c["z"],
                                                            ratio=ratio, randomized=c["randomized"], u_rand=c["u_rand"],
                                                            label_layout="reference")
    assert torch.allclose(mask, want["conf_pred_mask"], rtol=0, atol=1e-6), (mask - want["conf_pred_mask"]).abs().max()
    assert torch.allclose(ood, want["ood_mask"], rtol=RTOL, atol=1e-12)
    if c["optimistic"]:
        mu = torch.stack([gp.mean + gp.k(c["Xq"], gp.X) @ gp.alpha for gp in gps], -1)      # (b, q, t)
        assert torch.equal((grid >= mu.unsqueeze(1)).double(), want["opt_mask"].double())
    else:
        assert bool((want["opt_mask"] == 1).all())


def _grad_close(got, want, rtol=1e-6):
    scale = want.abs().max().clamp_min(1e-300)
    return torch.allclose(got, want, rtol=rtol, atol=float(1e-8 * scale))


@pytest.mark.parametrize("idx", range(6))
def test_oracle_gradients_against_reference_source_golden(idx):
    """d acq / d X: torch.autograd through the reference's OWN classes (over the exact-GP stand-in; what
    conformalbo/optim/optimize.py:224-235 does every SGLD step) recorded in tests/golden/ref_path_grad.pt -- cUCB q = 1, 2,
    cEI with two MC samples, cNEI q = 2, a pending point.

    * q = 1: autograd through both oracles equals the reference's gradient (1e-6; measured 1e-14).
    * joint batches (q + pending >= 2): the reference builds the grid under torch.no_grad() (helpers.py:118-119), but the
      predictive covariance is LAZY and its prior block k(x_j, x_k) is only evaluated when the grid is sampled, outside the
      no_grad block: the grid leaks a gradient through k(x_j, x_k).  `dense.conformal_acquisition(lazy_prior_grad=True)`
      reproduces the reference's gradient to 1e-6 (measured 1e-14) -- the diagnosis, quantified; the default of the oracles and
      the CUDA reverse pass treat the grid as the constant the no_grad block intends, which differs by O(1)."""
    c = acq_cases("ref_path_grad.pt")[idx]
    gp, Xq, kw = oracle_inputs(c)
    want = c["out"]
    joint = Xq.shape[1] >= 2
    assert want["dacq_dX"].abs().max() > 0

    def grad_of(mod, **extra):
        X = Xq.clone().requires_grad_(True)
        got = mod.conformal_acquisition(gp, X, **kw, **extra)
        assert torch.allclose(got["acq"].detach(), want["acq"], rtol=RTOL, atol=1e-9)
        (g,) = torch.autograd.grad(got["acq"].sum(), X)
        return g[:, :c["q"]]                                       # pending points carry no gradient

    big = c["N"] > 256                                             # N = 1024: closed form only (the dense refit needs GBs under autograd)
    if not big:
        assert _grad_close(grad_of(dense, lazy_prior_grad=True), want["dacq_dX"])
    assert _grad_close(grad_of(closed_form, lazy_prior_grad=True), want["dacq_dX"])
    g_const = grad_of(closed_form)
    if not big:
        assert _grad_close(g_const, grad_of(dense))                # the two oracles share ONE convention
    if not joint:
        assert _grad_close(g_const, want["dacq_dX"])
    else:
        rel = float((g_const - want["dacq_dX"]).abs().max() / want["dacq_dX"].abs().max())
        assert rel > 1e-3                                          # the leak is not a rounding effect


@pytest.mark.parametrize("idx", range(8))
def test_closed_form_against_more_reference_source_goldens(idx):
    """tests/golden/ref_path_more.pt (oracle/make_golden_path.py --more): cEI q = 2 with two MC samples, q = 4, cNEI with a
    pending point, the hard path with two refinements, q = 3 non-optimistic with refinement, G = 128, three MC samples
    without the optimistic mask, temp = 1e-6 with q = 2 and refinement."""
    c = acq_cases("ref_path_more.pt")[idx]
    gp, Xq, kw = oracle_inputs(c)
    check_against_golden(c, closed_form.conformal_acquisition(gp, Xq, **kw))
