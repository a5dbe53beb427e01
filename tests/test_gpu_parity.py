"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle.

Tolerance (BASELINE.json north_star): acquisition values and band endpoints within 1e-6 relative
in fp64; masks identical except within that tolerance of a rank tie (soft masks: 1e-6 absolute).
"""
import math

import pytest
import torch

from oracle import closed_form, dense, pure
from oracle import gp as ogp
from tests.problems import make_inputs, make_problem, oracle_state

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _dev():
    assert torch.cuda.is_available(), "run with -m gpu on a CUDA box"
    return torch.device("cuda:0")


def _state(p, dev):
    from conformal_bayesopt_b200 import GPModelState
    return GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)


def _ratio(d, dev):
    from conformal_bayesopt_b200 import ops
    r = ogp.make_ratio_mlp(d)
    return r, ops.RatioWeights(*[t.to(dev) for t in (r.W1, r.b1, r.W2, r.b2, r.W3, r.b3)])


def _run(p, inp, dev, ratio_w=None, outputs=(), **kw):
    from conformal_bayesopt_b200 import ops
    out = ops.conformal_acq(_state(p, dev), inp["Xq"].to(dev), eps_grid=inp["eps_grid"][:, 0].to(dev), z=inp["z"].to(dev),
                            eps_mc=inp["eps_mc"][:, :, 0].to(dev), u_rand=inp["u_rand"][:, :, 0].to(dev), ratio=ratio_w,
                            outputs=outputs, **kw)
    torch.cuda.synchronize()
    return {k: v.cpu() for k, v in out.items()}


def _close(a, b, rtol=RTOL, atol=1e-9):
    return torch.allclose(a, b, rtol=rtol, atol=atol)


def test_library_loaded_is_in_tree():
    from conformal_bayesopt_b200 import _lib
    lib = _lib.load()
    assert lib.cbo_version() == 104
    assert "conformal_bayesopt_b200/lib/libcbo.so" in _lib.lib_path()


@pytest.mark.parametrize("M,K,tri", [(100, 256, 1), (300, 384, 2), (129, 128, 0), (1, 256, 1)])
def test_dmma_gemm_matches_torch(M, K, tri):
    from conformal_bayesopt_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(M)
    A = torch.randn(M, K, generator=g, dtype=torch.float64).to(dev)
    B = torch.randn(K, K, generator=g, dtype=torch.float64).to(dev)
    B = torch.tril(B) if tri == 1 else (torch.triu(B) if tri == 2 else B)
    C, rn = ops.gemm_tn(A, B, tri, rownorm=True)
    ref = A @ B.T
    assert torch.allclose(C, ref, rtol=1e-12, atol=1e-11)
    assert torch.allclose(rn.sum(1), (ref * ref).sum(1), rtol=1e-11)


def test_kernel_rows_and_posterior_mean():
    from conformal_bayesopt_b200 import ops
    dev = _dev()
    p = make_problem(150, 5, seed=2)
    gp = oracle_state(p)
    Xq = torch.rand(77, 5, dtype=torch.float64)
    U, mu = ops.kernel_rows(_state(p, dev), Xq.to(dev))
    Uo = gp.k(Xq, gp.X)
    assert torch.allclose(U.cpu()[:, :150], Uo, rtol=1e-12, atol=1e-14)
    assert torch.all(U.cpu()[:, 150:] == 0)
    assert torch.allclose(mu.cpu(), gp.mean + Uo @ gp.alpha, rtol=1e-10, atol=1e-12)


CASES = [
    # N, d, b, G, temp, randomized, ratio, refine, kind, S
    (40, 3, 6, 16, 3e-2, False, False, 0, "ucb", 1),
    (200, 8, 37, 64, 3e-2, True, True, 0, "ucb", 1),
    (130, 4, 9, 32, 0.0, True, True, 0, "ucb", 2),
    (64, 2, 20, 64, 1e-6, True, False, 2, "ei", 1),
    (300, 8, 50, 64, 3e-2, True, True, 3, "ucb", 3),
    (8, 2, 128, 64, 3e-2, True, False, 0, "ucb", 1),       # BASELINE config 1 shape: branin-like 2-D, 8 points, 128 raw samples
    (32, 12, 224, 64, 1e-6, True, True, 1, "ucb", 1),      # BASELINE config 4 shape: tabular pool b=224, d=12, N=32
    (257, 8, 5, 7, 1e-2, False, True, 0, "ei", 4),         # ragged: N, G not multiples of anything
    (128, 3, 1, 2, 3e-2, False, False, 0, "ucb", 1),       # minimum grid, single candidate
]


@pytest.mark.parametrize("N,d,b,G,temp,rand,ratio_on,refine,kind,S", CASES)
def test_conformal_acq_matches_closed_form_oracle(N, d, b, G, temp, rand, ratio_on, refine, kind, S):
    dev = _dev()
    p = make_problem(N, d, seed=N)
    inp = make_inputs(N, d, b, 1, G, S=S, seed=N)
    ratio, ratio_w = _ratio(d, dev) if ratio_on else (None, None)
    kw = dict(kind=kind, alpha=0.1, temp=temp, grid_res=G, max_grid_refinements=refine, randomized=rand,
              optimistic=True, beta_prime=math.sqrt(0.2 * math.pi / 2), best_f=0.3)
    want = closed_form.conformal_acquisition(oracle_state(p), inp["Xq"], eps_grid=inp["eps_grid"], z=inp["z"],
                                             eps_mc=inp["eps_mc"], ratio=ratio, u_rand=inp["u_rand"], **kw)
    got = _run(p, inp, dev, ratio_w, outputs=("target_grid", "grid_logp", "conf_pred_mask", "ood_mask", "post_mean", "post_var"), **kw)
    assert _close(got["acq"], want["acq"])
    assert _close(got["target_grid"], want["target_grid"][:, :, 0])
    assert _close(got["grid_logp"], want["logp"])
    assert _close(got["post_mean"], want["mu"][:, 0]) and _close(got["post_var"], want["S"][:, 0, 0])
    assert torch.allclose(got["conf_pred_mask"], want["mask"][:, :, 0], rtol=0, atol=1e-6)
    assert torch.allclose(got["ood_mask"], want["ood"][:, 0, 0], rtol=RTOL, atol=1e-12)
    # band endpoints through the CUDA conf_mask_to_bounds
    from conformal_bayesopt_b200 import ops
    lb, ub = ops.conf_mask_to_bounds(got["target_grid"].to(dev)[:, :, None, None], got["conf_pred_mask"].to(dev)[:, :, None, None])
    lbo, ubo = pure.conf_mask_to_bounds(want["target_grid"].unsqueeze(-1), want["mask"].unsqueeze(-1))
    fin = torch.isfinite(lbo)
    assert torch.equal(torch.isfinite(lb.cpu()), fin)
    if temp > 1e-3:   # interpolated endpoints are smooth in the mask only on the soft path
        assert _close(lb.cpu()[fin], lbo[fin], atol=1e-7) and _close(ub.cpu()[fin], ubo[fin], atol=1e-7)


def test_conformal_acq_matches_dense_literal_oracle():
    """Directly against the reference-algorithm restatement (dense (N+1)^2 refit per (cand, grid))."""
    dev = _dev()
    N, d, b, G = 96, 4, 12, 24
    p = make_problem(N, d, seed=5)
    inp = make_inputs(N, d, b, 1, G, S=2, seed=5)
    ratio, ratio_w = _ratio(d, dev)
    kw = dict(kind="ucb", alpha=0.1, temp=3e-2, grid_res=G, max_grid_refinements=1, randomized=True, optimistic=True,
              beta_prime=1.1)
    want = dense.conformal_acquisition(oracle_state(p), inp["Xq"], eps_grid=inp["eps_grid"], z=inp["z"],
                                       eps_mc=inp["eps_mc"], ratio=ratio, u_rand=inp["u_rand"], **kw)
    got = _run(p, inp, dev, ratio_w, outputs=("conf_pred_mask",), **kw)
    assert _close(got["acq"], want["acq"])
    assert torch.allclose(got["conf_pred_mask"], want["mask"][:, :, 0], rtol=0, atol=1e-6)


def test_mask_to_bounds_against_reference_golden(golden_dir):
    import os
    from conformal_bayesopt_b200 import ops
    dev = _dev()
    gold = torch.load(os.path.join(golden_dir, "ref_pure_functions.pt"), weights_only=False)
    for case in gold["conf_mask_to_bounds"]:
        lb, ub = ops.conf_mask_to_bounds(case["grid"].double().to(dev), case["mask"].double().to(dev))
        wl, wu = case["lb"].double(), case["ub"].double()
        fin = torch.isfinite(wl)
        assert torch.equal(torch.isfinite(lb.cpu()), fin)
        assert torch.allclose(lb.cpu()[fin], wl[fin], rtol=1e-6, atol=1e-7)
        assert torch.allclose(ub.cpu()[fin], wu[fin], rtol=1e-6, atol=1e-7)


def test_ratio_forward_against_reference_golden(golden_dir):
    import os
    from conformal_bayesopt_b200 import ops
    dev = _dev()
    gold = torch.load(os.path.join(golden_dir, "ref_pure_functions.pt"), weights_only=False)
    for c in gold["ratio_forward"]:
        sd = c["state_dict"]
        w = ops.RatioWeights(*[sd[k].to(dev) for k in ("0.weight", "0.bias", "2.weight", "2.bias", "4.weight", "4.bias")])
        out = ops.ratio_forward(w, c["x"].to(dev))
        assert torch.allclose(out.cpu(), c["out"], rtol=1e-9, atol=0)


def test_full_size_properties():
    """BASELINE config 2 size (N=1024, d=8, G=64, b=4096): size-independent properties + an oracle spot check."""
    from conformal_bayesopt_b200 import ops
    dev = _dev()
    N, d, b, G = 1024, 8, 4096, 64
    p = make_problem(N, d, seed=0)
    inp = make_inputs(N, d, b, 1, G, seed=0)
    kw = dict(kind="ucb", alpha=1 / math.sqrt(N), temp=3e-2, grid_res=G, randomized=False, optimistic=True, beta_prime=1.0)
    outs = ("conf_pred_mask", "target_grid", "post_mean", "post_var")
    got = _run(p, inp, dev, outputs=outs, **kw)
    assert torch.isfinite(got["acq"]).all()
    assert (got["post_var"] > p["noise"] * (1 - 1e-9)).all()              # Schur complement >= noise
    assert ((got["conf_pred_mask"] >= 0) & (got["conf_pred_mask"] <= 1)).all()
    # chunking invariance: different internal chunk sizes give bit-identical results
    got2 = _run(p, inp, dev, outputs=outs, chunk=1000, **kw)
    assert torch.equal(got["acq"], got2["acq"]) and torch.equal(got["conf_pred_mask"], got2["conf_pred_mask"])
    # candidate-order invariance: a permuted batch gives permuted results, bit for bit
    perm = torch.randperm(b, generator=torch.Generator().manual_seed(1))
    inp_p = dict(inp, Xq=inp["Xq"][perm], u_rand=inp["u_rand"][perm])
    got3 = _run(p, inp_p, dev, **kw)
    assert torch.equal(got3["acq"], got["acq"][perm])
    # oracle spot check on 48 candidates of the full-size problem
    sel = torch.arange(0, b, b // 48)[:48]
    want = closed_form.conformal_acquisition(oracle_state(p), inp["Xq"][sel], eps_grid=inp["eps_grid"], z=inp["z"],
                                             eps_mc=inp["eps_mc"], **kw)
    assert _close(got["acq"][sel], want["acq"])
    assert torch.allclose(got["conf_pred_mask"][sel], want["mask"][:, :, 0], rtol=0, atol=1e-6)


def test_train_permutation_invariance():
    """permuting the training points (and the probe rows with them) leaves the outputs unchanged."""
    dev = _dev()
    N, d, b, G = 100, 3, 10, 32
    p = make_problem(N, d, seed=9)
    inp = make_inputs(N, d, b, 1, G, seed=9)
    kw = dict(kind="ucb", alpha=0.1, temp=3e-2, grid_res=G, randomized=False, optimistic=True, beta_prime=1.0)
    a = _run(p, inp, dev, **kw)
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(3))
    p2 = dict(p, X=p["X"][perm], y=p["y"][perm])
    z2 = torch.cat([inp["z"][:N][perm], inp["z"][N:]])
    b_ = _run(p2, dict(inp, z=z2), dev, **kw)
    assert torch.allclose(a["acq"], b_["acq"], rtol=1e-8, atol=1e-10)


def test_error_conventions():
    from conformal_bayesopt_b200 import _lib, ops
    dev = _dev()
    p = make_problem(16, 2, seed=1)
    inp = make_inputs(16, 2, 4, 1, 8, seed=1)
    with pytest.raises(_lib.CboError):       # grid_res >= 2 (helpers.py:113)
        _run(p, dict(inp, eps_grid=inp["eps_grid"][:1], eps_mc=inp["eps_mc"][:, :1]), dev, kind="ucb", alpha=0.1, temp=0.1, grid_res=1)
    with pytest.raises(_lib.CboError):       # no CPU fallback
        ops.conformal_acq(_state(p, dev), inp["Xq"], eps_grid=inp["eps_grid"][:, 0], z=inp["z"], eps_mc=inp["eps_mc"][:, :, 0],
                          kind="ucb", alpha=0.1, temp=0.1, grid_res=8)
    # NaN candidates propagate to NaN values, finite candidates are unaffected
    Xq = inp["Xq"].clone()
    Xq[1] = float("nan")
    got = _run(p, dict(inp, Xq=Xq), dev, kind="ucb", alpha=0.1, temp=0.1, grid_res=8, beta_prime=1.0)
    assert torch.isnan(got["acq"][1]) and torch.isfinite(got["acq"][[0, 2, 3]]).all()


QCASES = [
    # N, d, q, b, G, temp, randomized, ratio, refine, kind, S
    (33, 2, 2, 5, 12, 3e-2, True, True, 0, "ucb", 2),
    (33, 2, 2, 5, 12, 0.0, False, False, 1, "ucb", 1),
    (20, 4, 3, 4, 8, 1e-2, True, True, 3, "ei", 2),
    (8, 2, 2, 128, 64, 3e-2, True, False, 0, "ucb", 1),     # BASELINE config 1: branin-like, N=8, q_batch_size=2, 128 raw samples
    (150, 8, 4, 9, 32, 3e-2, True, True, 2, "ucb", 1),
    (130, 5, 2, 300, 64, 1e-6, True, True, 1, "ei", 1),
]


@pytest.mark.parametrize("N,d,q,b,G,temp,rand,ratio_on,refine,kind,S", QCASES)
def test_joint_batches_match_oracle(N, d, q, b, G, temp, rand, ratio_on, refine, kind, S):
    """q > 1 against the closed-form oracle (itself pinned against the dense refit for q = 2, 3)."""
    from conformal_bayesopt_b200 import ops
    dev = _dev()
    p = make_problem(N, d, seed=N + q)
    inp = make_inputs(N, d, b, q, G, S=S, seed=N + q)
    ratio, ratio_w = _ratio(d, dev) if ratio_on else (None, None)
    kw = dict(kind=kind, alpha=0.2, temp=temp, grid_res=G, max_grid_refinements=refine, randomized=rand,
              optimistic=True, beta_prime=math.sqrt(0.2 * math.pi / 2), best_f=0.3)
    want = closed_form.conformal_acquisition(oracle_state(p), inp["Xq"], eps_grid=inp["eps_grid"], z=inp["z"],
                                             eps_mc=inp["eps_mc"], ratio=ratio, u_rand=inp["u_rand"], **kw)
    got = ops.conformal_acq(_state(p, dev), inp["Xq"].to(dev), eps_grid=inp["eps_grid"].to(dev), z=inp["z"].to(dev),
                            eps_mc=inp["eps_mc"].to(dev), u_rand=inp["u_rand"].to(dev), ratio=ratio_w,
                            outputs=("target_grid", "grid_logp", "conf_pred_mask", "ood_mask", "post_mean", "post_var"), **kw)
    torch.cuda.synchronize()
    got = {k: v.cpu() for k, v in got.items()}
    assert _close(got["post_mean"], want["mu"]) and _close(got["post_var"], want["S"])
    # q >= 3 with refinement: the rescale factor |diff|^(2/q) is a float32 pow in the reference (helpers.py:167); device and
    # host float pow may differ by one float ulp, i.e. 6e-8 relative on (y_g - mu) -- allow for it in absolute terms
    gtol = 1e-7 if (q >= 3 and refine > 0) else 1e-9
    assert _close(got["target_grid"], want["target_grid"], atol=gtol)
    assert _close(got["grid_logp"], want["logp"])
    assert torch.allclose(got["conf_pred_mask"], want["mask"], rtol=0, atol=1e-6)
    assert torch.allclose(got["ood_mask"], want["ood"][:, 0, :], rtol=RTOL, atol=1e-12)
    assert _close(got["acq"], want["acq"])


BIGCASES = [
    # N, d, b, G, temp, S, Sp  -- shapes that leave the tuned configuration (fewer warp pairs per SM, unstaged rows kernel,
    # 16 probes, many MC samples, long grids)
    (2500, 3, 6, 16, 3e-2, 1, 8),
    (1300, 33, 5, 200, 1e-2, 16, 16),
    (4000, 2, 3, 8, 3e-2, 2, 4),
    (600, 64, 4, 1000, 3e-2, 1, 8),
]


@pytest.mark.parametrize("N,d,b,G,temp,S,Sp", BIGCASES)
def test_shapes_outside_the_tuned_configuration(N, d, b, G, temp, S, Sp):
    from conformal_bayesopt_b200 import ops
    dev = _dev()
    p = make_problem(N, d, seed=N, fn="sin", lengthscale=0.5 * math.sqrt(d / 3.0))
    inp = make_inputs(N, d, b, 1, G, S=S, Sp=Sp, seed=N)
    ratio, ratio_w = _ratio(d, dev)
    kw = dict(kind="ucb", alpha=0.1, temp=temp, grid_res=G, max_grid_refinements=1, randomized=True, optimistic=True,
              beta_prime=1.3)
    want = closed_form.conformal_acquisition(oracle_state(p), inp["Xq"], eps_grid=inp["eps_grid"], z=inp["z"],
                                             eps_mc=inp["eps_mc"], ratio=ratio, u_rand=inp["u_rand"], **kw)
    got = _run(p, inp, dev, ratio_w, outputs=("conf_pred_mask", "post_var"), **kw)
    assert _close(got["post_var"], want["S"][:, 0, 0])
    assert torch.allclose(got["conf_pred_mask"], want["mask"][:, :, 0], rtol=0, atol=1e-6)
    assert _close(got["acq"], want["acq"])


def test_unsupported_shapes_are_refused():
    from conformal_bayesopt_b200 import _lib
    dev = _dev()
    p = make_problem(16, 2, seed=1)
    inp = make_inputs(16, 2, 4, 1, 8, S=17, seed=1)               # 17 MC samples > MAX_S
    with pytest.raises(_lib.CboError, match="unsupported"):
        _run(p, inp, dev, kind="ucb", alpha=0.1, temp=0.1, grid_res=8, beta_prime=1.0)
    inp = make_inputs(16, 2, 4, 5, 8, seed=1)                      # q = 5
    with pytest.raises(_lib.CboError):
        from conformal_bayesopt_b200 import ops
        ops.conformal_acq(_state(p, dev), inp["Xq"].to(dev), eps_grid=inp["eps_grid"].to(dev), z=inp["z"].to(dev),
                          eps_mc=inp["eps_mc"].to(dev), kind="ucb", alpha=0.1, temp=0.1, grid_res=8, beta_prime=1.0)
