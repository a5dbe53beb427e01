"""GPU: the CUDA path (C ABI and the drop-in classes) against goldens produced by EXECUTING the reference's own
conformalbo/helpers.py and conformalbo/acquisitions.py (oracle/make_golden_path.py; reference source unmodified, only
the gpytorch/botorch exact-GP algebra comes from the stand-ins in oracle/_shim).

North-star tolerances: acquisition values, grid, band endpoints 1e-6 relative; soft masks 1e-6 absolute; hard masks and
the selected-candidate (arg-max) index identical."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-6
DEV = "cuda:0"


def _cases(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def _scalar(v):
    """python float of a scalar hyper-parameter (NOT via torch.as_tensor(float): that would round it to float32)."""
    return float(v.reshape(-1)[0]) if torch.is_tensor(v) else float(v)


def _state(p):
    from conformal_bayesopt_b200 import GPModelState
    return GPModelState.fit_fixed(p["X"], p["y"], _scalar(p["mean"]), p["lengthscale"], _scalar(p["outputscale"]),
                                  _scalar(p["noise"]), device=DEV)


def _ratio(c):
    from conformal_bayesopt_b200 import ops
    return ops.RatioWeights(*[w.to(DEV) for w in c["ratio_weights"]]) if c["ratio"] else None


def _check(c, acq, out):
    want = c["out"]
    b, G, q = c["b"], c["G"], c["q"] + c["x_pending"]
    acq = acq.cpu()
    assert torch.allclose(acq, want["acq"], rtol=RTOL, atol=1e-9), float((acq - want["acq"]).abs().max())
    assert int(acq.argmax()) == want["argmax"]                       # selected-candidate index (utils.py:283)
    o = {k: v.cpu().reshape(b, G, q) if k in ("target_grid", "conf_pred_mask") else v.cpu() for k, v in out.items()}
    assert torch.allclose(o["target_grid"], want["target_grid"][..., 0], rtol=RTOL, atol=1e-9)
    assert torch.allclose(o["grid_logp"].reshape(b, G), want["grid_logp"], rtol=RTOL, atol=1e-9)
    wm = want["conf_pred_mask"][..., 0]
    if c["temp"] > 0:
        assert torch.allclose(o["conf_pred_mask"], wm, rtol=0, atol=1e-6), float((o["conf_pred_mask"] - wm).abs().max())
    else:
        assert torch.equal(o["conf_pred_mask"], wm)
    assert torch.allclose(o["ood_mask"].reshape(b, 1, q).expand(b, G, q), want["ood_mask"][..., 0], rtol=RTOL, atol=1e-12)


def _call_abi(c):
    from conformal_bayesopt_b200 import ops
    st = _state(c["problem"])
    Xq = c["Xq"] if c["X_pending"] is None else torch.cat([c["Xq"], c["X_pending"].expand(c["b"], -1, -1)], dim=1)
    kw = dict(kind=c["kind"], alpha=c["alpha"], temp=c["temp"], grid_res=c["G"], max_grid_refinements=c["refine"],
              randomized=c["randomized"], optimistic=c["optimistic"], beta_prime=math.sqrt(c["beta"] * math.pi / 2),
              best_f=c["best_f"])
    if c["kind"] == "nei":
        kw["nei_baseline"] = ops.NeiBaseline.build(st, c["X_baseline"].to(DEV))
    out = ops.conformal_acq(st, Xq.to(DEV), eps_grid=c["eps_grid"].to(DEV), z=c["z"].to(DEV), eps_mc=c["eps_mc"].to(DEV),
                            u_rand=None if c["u_rand"] is None else c["u_rand"].to(DEV), ratio=_ratio(c),
                            outputs=("target_grid", "grid_logp", "conf_pred_mask", "ood_mask"), **kw)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("idx", range(12))
def test_cuda_abi_against_reference_source_golden(idx):
    c = _cases("ref_path.pt")["acq"][idx]
    out = _call_abi(c)
    _check(c, out["acq"], out)
    if "lb" in c["out"]:
        from conformal_bayesopt_b200 import ops
        b, G = c["b"], c["G"]
        lb, ub = ops.conf_mask_to_bounds(out["target_grid"].reshape(b, G, 1, 1), out["conf_pred_mask"].reshape(b, G, 1, 1))
        lb, ub = lb.cpu().reshape(-1), ub.cpu().reshape(-1)
        fin = torch.isfinite(c["out"]["lb"])
        assert torch.equal(torch.isfinite(lb), fin)
        if c["temp"] > 1e-3:
            assert torch.allclose(lb[fin], c["out"]["lb"][fin], rtol=RTOL, atol=1e-7)
            assert torch.allclose(ub[fin], c["out"]["ub"][fin], rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("idx", range(3))
def test_cuda_abi_against_reference_source_golden_n1024(idx):
    """BASELINE configs[1] sizes: N=1024, d=8, G=64 (cUCB; cEI + ratio + randomised; q=2)."""
    c = _cases("ref_path_n1024.pt")["acq"][idx]
    out = _call_abi(c)
    _check(c, out["acq"], out)


@pytest.mark.parametrize("name", ["cfg1_ucb_q2", "cfg4_pool", "small_ei_refine3", "small_nei", "small_pending"])
def test_acquisition_classes_against_reference_source_golden(name):
    """Same goldens through the drop-in classes (constructor keywords as the reference's hydra configs pass them)."""
    import conformal_bayesopt_b200 as cbo
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    c = [x for x in _cases("ref_path.pt")["acq"] if x["name"] == name][0]
    st = _state(c["problem"])
    G, S = c["G"], c["S"]
    gs, s = IIDNormalSampler(G), IIDNormalSampler(S, batch_range=(0, -3))
    gs.set_base_samples(c["eps_grid"])
    s.set_base_samples(c["eps_mc"])
    kw = dict(alpha=c["alpha"], temp=c["temp"], grid_res=G, max_grid_refinements=c["refine"], ratio_estimator=_ratio(c),
              optimistic=c["optimistic"], grid_sampler=gs, randomized=c["randomized"], model=st, sampler=s,
              X_pending=None if c["X_pending"] is None else c["X_pending"].to(DEV))
    if c["kind"] == "ucb":
        acq_fn = cbo.qConformalUpperConfidenceBound(beta=c["beta"], **kw)
    elif c["kind"] == "ei":
        acq_fn = cbo.qConformalExpectedImprovement(best_f=c["best_f"], **kw)
    else:
        acq_fn = cbo.qConformalNoisyExpectedImprovement(X_baseline=c["X_baseline"].to(DEV), **kw)
    acq_fn.fixed_probes = c["z"]
    acq_fn.fixed_uniforms = None if c["u_rand"] is None else c["u_rand"].to(DEV)
    with torch.no_grad():
        acq = acq_fn(c["Xq"].to(DEV))
    assert acq.shape == (c["b"],)
    assert torch.allclose(acq.cpu(), c["out"]["acq"], rtol=RTOL, atol=1e-9)
    assert int(acq.argmax()) == c["out"]["argmax"]


@pytest.mark.parametrize("idx", range(2))
def test_assess_coverage_against_reference_source_golden(idx):
    """helpers.assess_coverage (credible vs conformal coverage of a hold-out set, per-point base samples)."""
    from conformal_bayesopt_b200 import helpers
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    c = [x for x in _cases("ref_path.pt")["coverage"] if x["t"] == 1][idx]
    st = _state(c["problem"])
    G, n = c["G"], c["n"]
    gs = IIDNormalSampler(G, resample=False, collapse_batch_dims=False)
    gs.set_base_samples(c["eps_grid"].reshape(G, n, 1))
    u = None if c["u_rand"] is None else c["u_rand"].reshape(n, G).to(DEV)
    std_cov, conf_cov = helpers.assess_coverage(st, c["X_holdout"].to(DEV), c["Y_holdout"].to(DEV), c["alpha"], c["temp"], G,
                                                c["refine"], _ratio(c), c["randomized"], probes=c["z"], u_rand=u,
                                                grid_sampler=gs)
    assert abs(std_cov - c["out"]["std_coverage"]) < 1e-6
    assert abs(conf_cov - c["out"]["conformal_coverage"]) < 1e-6
    tg, _, mk, _, _ = helpers.construct_conformal_bands(st, c["X_holdout"].to(DEV)[:, None], c["alpha"], c["temp"], G,
                                                        c["refine"], gs, _ratio(c), randomized=c["randomized"],
                                                        probes=c["z"], u_rand=u)
    assert torch.allclose(tg.cpu(), c["out"]["target_grid"], rtol=RTOL, atol=1e-9)
    assert torch.allclose(mk.cpu(), c["out"]["conf_pred_mask"], rtol=0, atol=1e-6)
    lb, ub = helpers.conf_mask_to_bounds(tg, mk)
    fin = torch.isfinite(c["out"]["lb"])
    assert torch.equal(torch.isfinite(lb.cpu()), fin)
    if c["temp"] > 1e-3:
        assert torch.allclose(lb.cpu()[fin], c["out"]["lb"][fin], rtol=RTOL, atol=1e-7)
        assert torch.allclose(ub.cpu()[fin], c["out"]["ub"][fin], rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("name", ["cov_t2", "cov_t2_rand"])
def test_multi_output_bands_and_coverage_intent_layout(name):
    """t = 2 (helpers.py:252-258, 336-387, 408-411): the device path implements label_layout='intent' with the reference's
    joint refinement decisions; checked against oracle/multi.py (whose 'reference' layout is pinned to the reference-source
    goldens in tests/test_oracle_ref_path.py).  label_layout='reference' is refused loudly on the device."""
    from conformal_bayesopt_b200 import GPModelState, helpers
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    from oracle import gp as ogp
    from oracle import multi
    c = [x for x in _cases("ref_path.pt")["coverage"] if x["name"] == name][0]
    p, t, G, n = c["problem"], c["t"], c["G"], c["n"]
    hyp = [(float(p["mean"][k]), p["lengthscale"][k], float(p["outputscale"][k]), float(p["noise"][k])) for k in range(t)]
    gps = [ogp.build_state(p["X"], p["Y"][:, k], *hyp[k]) for k in range(t)]
    states = [GPModelState.fit_fixed(p["X"], p["Y"][:, k], *hyp[k], device=DEV) for k in range(t)]
    oratio = ogp.RatioMLP(*c["ratio_weights"]) if c["ratio"] else None
    want = multi.assess_coverage(gps, c["X_holdout"], c["Y_holdout"], c["alpha"], c["temp"], G, c["refine"], c["eps_grid"],
                                 c["z"], ratio=oratio, randomized=c["randomized"], u_rand=c["u_rand"], label_layout="intent")
    gs = IIDNormalSampler(G, resample=False, collapse_batch_dims=False)
    gs.set_base_samples(c["eps_grid"])                                     # (G, n, 1, t)
    u = None if c["u_rand"] is None else c["u_rand"].to(DEV)
    std_cov, conf_cov = helpers.assess_coverage(states, c["X_holdout"].to(DEV), c["Y_holdout"].to(DEV), c["alpha"], c["temp"], G,
                                                c["refine"], _ratio(c), c["randomized"], probes=c["z"], u_rand=u, grid_sampler=gs)
    assert abs(std_cov - want[0]) < 1e-6 and abs(conf_cov - want[1]) < 1e-6
    tg, w, mk, _, ood = helpers.construct_conformal_bands(states, c["X_holdout"].to(DEV)[:, None], c["alpha"], c["temp"], G,
                                                          c["refine"], gs, _ratio(c), randomized=c["randomized"],
                                                          probes=c["z"], u_rand=u)
    assert tg.shape == (n, G, 1, t) and w.shape == (n, G, 1, 1) and mk.shape == (n, G, 1, t) and ood.shape == (n, G, 1, t)
    assert torch.allclose(tg.cpu(), want[2], rtol=RTOL, atol=1e-9)
    assert torch.allclose(mk.cpu(), want[3], rtol=0, atol=1e-6)
    lb, ub = helpers.conf_mask_to_bounds(tg, mk)
    fin = torch.isfinite(want[4])
    assert torch.equal(torch.isfinite(lb.cpu()), fin)
    assert torch.allclose(lb.cpu()[fin], want[4][fin], rtol=RTOL, atol=1e-7) and torch.allclose(ub.cpu()[fin], want[5][fin], rtol=RTOL, atol=1e-7)
    with pytest.raises(NotImplementedError):
        helpers.construct_conformal_bands(states, c["X_holdout"].to(DEV)[:, None], c["alpha"], c["temp"], G, 0, gs, None,
                                          label_layout="reference")


@pytest.mark.parametrize("name", ["ehvi_q1", "ehvi_q2_rand"])
def test_cehvi_inner_expectation_against_reference_source_golden(name):
    """qConformalExpectedHypervolumeImprovement on a two-output model against what the reference's OWN `_conformal_fwd` hands
    to `_conformal_integration` (oracle/make_golden_path.py::run_ehvi_case): `values` (b, G) -- `_compute_qehvi` under every
    refit model, the cEHVI inner expectation named by the north star (`k_ehvi_q`) -- the joint grid log-density and the
    optimistic mask.  (The masks of that golden follow the reference's label scramble for t > 1, App. C-6: they are checked
    on the band-construction path, `helpers.construct_conformal_bands(label_layout="reference")`; the integration itself is
    unpinned, App. C-5.)"""
    from conformal_bayesopt_b200 import GPModelState
    from conformal_bayesopt_b200.mobo_acquisitions import qConformalExpectedHypervolumeImprovement
    from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    c = [x for x in _cases("ref_path_ehvi.pt")["ehvi"] if x["name"] == name][0]
    p, want = c["problem"], c["out"]
    b, q, G, S, t = c["b"], c["q"], c["G"], c["S"], c["t"]
    states = [GPModelState.fit_fixed(p["X"], p["Y"][:, k], float(p["mean"][k]), p["lengthscale"][k], float(p["outputscale"][k]),
                                     float(p["noise"][k]), device=DEV) for k in range(t)]
    est = None
    if c["ratio"]:
        est = RatioEstimator(in_size=c["d"], device=torch.device(DEV), dtype=torch.float64)
        lin = [m for m in est._target_network if isinstance(m, torch.nn.Linear)]
        with torch.no_grad():
            for k, layer in enumerate(lin):
                layer.weight.copy_(c["ratio_weights"][2 * k])
                layer.bias.copy_(c["ratio_weights"][2 * k + 1])
    # a non-interleaved MultitaskMultivariateNormal consumes (..., q, t) base samples as the flat task-major event (App. B)
    remap = lambda e: e.reshape(*e.shape[:-2], e.shape[-1], e.shape[-2]).transpose(-1, -2).contiguous()
    gs = IIDNormalSampler(G)
    gs.set_base_samples(remap(c["eps_grid"][:, 0]))
    mc = IIDNormalSampler(S, batch_range=(0, -3))
    mc.set_base_samples(remap(c["eps_mc"][:, 0]))
    acq = qConformalExpectedHypervolumeImprovement(alpha=c["alpha"], temp=c["temp"], grid_res=G, max_grid_refinements=0,
                                                   ratio_estimator=est, optimistic=c["optimistic"], grid_sampler=gs,
                                                   randomized=c["randomized"], model=states, ref_point=c["ref_point"],
                                                   partitioning=(c["cell_lower"], c["cell_upper"]), sampler=mc)
    acq.fixed_probes = c["z"]
    acq.fixed_uniforms = c["u_rand"].permute(0, 1, 3, 2).contiguous().to(DEV) if c["randomized"] else None
    with torch.no_grad():
        val = acq(c["Xq"].to(DEV))
    assert val.shape == (b,) and torch.isfinite(val).all()
    got = acq.last_outputs
    values = got["values"].cpu().reshape(b, G)
    assert torch.allclose(values, want["values"], rtol=RTOL, atol=1e-9), float((values - want["values"]).abs().max())
    logp = sum(o["grid_logp"].cpu().reshape(b, G) for o in got["per_output"])
    assert torch.allclose(logp, want["grid_logp"].reshape(b, G), rtol=RTOL, atol=1e-9)
    if c["optimistic"]:
        for k, o in enumerate(got["per_output"]):
            opt = (o["target_grid"].cpu().reshape(b, G, q) >= o["post_mean"].cpu().reshape(b, 1, q)).double()
            assert torch.equal(opt, want["opt_mask"][..., k].double())


@pytest.mark.parametrize("idx", range(6))
def test_cuda_gradients_against_reference_source_golden(idx):
    """d acq / d X of cbo_conformal_acq_grad_f64 against torch.autograd through the reference's OWN classes
    (tests/golden/ref_path_grad.pt, oracle/make_golden_path.py --grad).  q = 1 (cUCB with ratio estimator and randomised step,
    cEI with two MC samples): equal to the reference's gradient, 1e-6.  Joint batches (cUCB q = 2, cNEI q = 2, a pending
    point): the reference leaks a gradient through the lazily evaluated prior block of the grid covariance
    (tests/test_oracle_ref_path.py::test_oracle_gradients_against_reference_source_golden); the library implements the constant
    grid the reference's no_grad block intends and is compared with autograd through the oracle under that convention."""
    from conformal_bayesopt_b200 import ops
    from tests.test_oracle_ref_path import oracle_inputs
    from oracle import closed_form
    c = _cases("ref_path_grad.pt")["acq"][idx]
    st = _state(c["problem"])
    Xq = c["Xq"] if c["X_pending"] is None else torch.cat([c["Xq"], c["X_pending"].expand(c["b"], -1, -1)], dim=1)
    kw = dict(kind=c["kind"], alpha=c["alpha"], temp=c["temp"], grid_res=c["G"], max_grid_refinements=0,
              randomized=c["randomized"], optimistic=c["optimistic"], beta_prime=math.sqrt(c["beta"] * math.pi / 2),
              best_f=c["best_f"])
    if c["kind"] == "nei":
        kw["nei_baseline"] = ops.NeiBaseline.build(st, c["X_baseline"].to(DEV))
    out = ops.conformal_acq(st, Xq.to(DEV), eps_grid=c["eps_grid"].to(DEV), z=c["z"].to(DEV), eps_mc=c["eps_mc"].to(DEV),
                            u_rand=None if c["u_rand"] is None else c["u_rand"].to(DEV), ratio=_ratio(c), want_grad=True, **kw)
    torch.cuda.synchronize()
    want = c["out"]
    assert torch.allclose(out["acq"].cpu(), want["acq"], rtol=RTOL, atol=1e-9)
    g = out["dacq_dX"].cpu()[:, :c["q"]]
    if Xq.shape[1] == 1:
        ref_g = want["dacq_dX"]
    else:
        gp, Xo, okw = oracle_inputs(c)
        Xo = Xo.clone().requires_grad_(True)
        (ref_g,) = torch.autograd.grad(closed_form.conformal_acquisition(gp, Xo, **okw)["acq"].sum(), Xo)
        ref_g = ref_g[:, :c["q"]]
    scale = ref_g.abs().max()
    assert torch.allclose(g, ref_g, rtol=RTOL, atol=float(1e-8 * scale)), float((g - ref_g).abs().max() / scale)


@pytest.mark.parametrize("idx", range(8))
def test_cuda_abi_against_more_reference_source_goldens(idx):
    """tests/golden/ref_path_more.pt: further shapes and flag combinations executed by the reference's own classes (see
    tests/test_oracle_ref_path.py::test_closed_form_against_more_reference_source_goldens)."""
    c = _cases("ref_path_more.pt")["acq"][idx]
    out = _call_abi(c)
    _check(c, out["acq"], out)


@pytest.mark.parametrize("name", ["grad_ucb_q2", "grad_pending", "grad_ucb_q1", "grad_nei_q2", "grad_cfg2_ucb_q2"])
def test_cuda_gradients_with_the_reference_grid_leak(name):
    """cbo_params.grid_grad_leak = 1: the reverse pass adds the gradient the reference's autograd propagates through the
    lazily evaluated prior block k(x_j, x_k) of the grid covariance.  cUCB q = 2, a q = 1 + pending batch and cNEI q = 2 (k_nei_values_q_bwd adds the part through its sample values) then
    equal `torch.autograd` through the reference's OWN classes (tests/golden/ref_path_grad.pt) to 1e-6; q = 1 is unaffected.
    grad_cfg2_ucb_q2 is the reference's default SGLD shape at the BASELINE size (N = 1024, d = 8, G = 64, q = 2, ratio estimator,
    randomised step)."""
    from conformal_bayesopt_b200 import ops
    c = [x for x in _cases("ref_path_grad.pt")["acq"] if x["name"] == name][0]
    st = _state(c["problem"])
    Xq = c["Xq"] if c["X_pending"] is None else torch.cat([c["Xq"], c["X_pending"].expand(c["b"], -1, -1)], dim=1)
    kw = dict(kind=c["kind"], alpha=c["alpha"], temp=c["temp"], grid_res=c["G"], max_grid_refinements=0,
              randomized=c["randomized"], optimistic=c["optimistic"], beta_prime=math.sqrt(c["beta"] * math.pi / 2),
              best_f=c["best_f"])
    if c["kind"] == "nei":
        kw["nei_baseline"] = ops.NeiBaseline.build(st, c["X_baseline"].to(DEV))
    out = {}
    for leak in (False, True):
        out[leak] = ops.conformal_acq(st, Xq.to(DEV), eps_grid=c["eps_grid"].to(DEV), z=c["z"].to(DEV), eps_mc=c["eps_mc"].to(DEV),
                                      u_rand=None if c["u_rand"] is None else c["u_rand"].to(DEV), ratio=_ratio(c),
                                      want_grad=True, grid_grad_leak=leak, **kw)
    torch.cuda.synchronize()
    want = c["out"]["dacq_dX"]
    scale = want.abs().max()
    g = out[True]["dacq_dX"].cpu()[:, :c["q"]]
    assert torch.equal(out[True]["acq"], out[False]["acq"])
    assert torch.allclose(g, want, rtol=RTOL, atol=float(1e-8 * scale)), float((g - want).abs().max() / scale)
    g0 = out[False]["dacq_dX"].cpu()[:, :c["q"]]
    if Xq.shape[1] == 1:
        assert torch.equal(g0, g)
    else:
        assert float((g0 - want).abs().max() / scale) > 1e-3           # without the switch: the constant-grid convention


@pytest.mark.parametrize("name", ["grad_ucb_q2", "grad_pending"])
def test_class_backward_with_the_reference_grid_leak(name):
    """`acq.grid_grad_leak = True` on the drop-in class: `acq(X).sum().backward()` (what optim/optimize.py:224-235 does) then
    gives the gradient torch.autograd gives through the reference's own class, 1e-6; `value_and_grad` and the graph-replayed
    call (`acq.graphed`) give the same rows."""
    import conformal_bayesopt_b200 as cbo
    from conformal_bayesopt_b200.sampling import IIDNormalSampler
    c = [x for x in _cases("ref_path_grad.pt")["acq"] if x["name"] == name][0]
    st = _state(c["problem"])
    G, S = c["G"], c["S"]
    gs, s = IIDNormalSampler(G), IIDNormalSampler(S, batch_range=(0, -3))
    gs.set_base_samples(c["eps_grid"])
    s.set_base_samples(c["eps_mc"])
    acq_fn = cbo.qConformalUpperConfidenceBound(
        beta=c["beta"], alpha=c["alpha"], temp=c["temp"], grid_res=G, max_grid_refinements=0, ratio_estimator=_ratio(c),
        optimistic=c["optimistic"], grid_sampler=gs, randomized=c["randomized"], model=st, sampler=s,
        X_pending=None if c["X_pending"] is None else c["X_pending"].to(DEV))
    acq_fn.fixed_probes = c["z"]
    acq_fn.fixed_uniforms = None if c["u_rand"] is None else c["u_rand"].to(DEV)
    acq_fn.grid_grad_leak = True
    X = c["Xq"].to(DEV).requires_grad_(True)
    val = acq_fn(X)
    val.sum().backward()
    want = c["out"]
    scale = want["dacq_dX"].abs().max()
    assert torch.allclose(val.detach().cpu(), want["acq"], rtol=RTOL, atol=1e-9)
    assert torch.allclose(X.grad.cpu(), want["dacq_dX"], rtol=RTOL, atol=float(1e-8 * scale))
    v2, g2 = acq_fn.value_and_grad(c["Xq"].to(DEV))
    assert torch.equal(g2, X.grad) and torch.equal(v2, val.detach())
    if c["X_pending"] is None:
        graphed = acq_fn.graphed(c["b"], c["q"])
        out = graphed(c["Xq"].to(DEV), z=c["z"].to(DEV), u_rand=acq_fn.fixed_uniforms)
        torch.cuda.synchronize()
        assert torch.equal(out["dacq_dX"], X.grad.reshape(out["dacq_dX"].shape))
