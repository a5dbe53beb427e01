"""Training-side device steps through the C-ABI against outputs of the reference's own RatioEstimator and
SGLD classes (tests/golden/ref_training.pt) and the numpy oracle (oracle/training.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import training as T

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_training.pt")
DEV = "cuda:0"


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD)


def _cmp(ref, got, tol):
    ref = ref.numpy()
    got = got.detach().cpu().numpy()
    stride = 1 if ref.size == got.size else 37
    np.testing.assert_allclose(got[::stride], ref, rtol=0, atol=tol)


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_train_step_kernel_replays_the_reference_class(gold, idx):
    """cbo_ratio_train_step_f64 driven with the minibatches the reference drew: parameters and target network after every
    callback.  Absolute tolerance 1e-9 (a 1e-6th of one Adam step): m / (sqrt(v) + eps) amplifies summation-order
    rounding only where |g| ~ eps."""
    from conformal_bayesopt_b200 import ops
    case = gold["ratio_training"][idx]
    d = case["d"]
    theta = case["theta0"].to(DEV).clone()
    m, v, target = torch.zeros_like(theta), torch.zeros_like(theta), torch.zeros_like(theta)
    loss = torch.zeros((), dtype=torch.float64, device=DEV)
    neg = case["neg"].to(DEV)
    pool = torch.zeros(0, d, dtype=torch.float64, device=DEV)
    step = 0
    for k, (xk, perm, trained) in enumerate(zip(case["callbacks"], case["perms"], case["trained"])):
        pool = torch.cat([pool, xk.to(DEV)])
        if trained:
            step += 1
            th_before = theta.cpu().numpy().copy()
            ops.ratio_train_step(theta, m, v, target, x_neg=neg, x_pos=pool, pos_idx=perm.to(DEV), n_pos=neg.shape[0],
                                 step=step, lr=case["lr"], weight_decay=case["weight_decay"],
                                 ema_weight=case["ema_weight"], loss=loss)
            x = np.concatenate([neg.cpu().numpy(), pool.cpu().numpy()[perm.numpy()]])
            y = np.concatenate([np.zeros(neg.shape[0]), np.ones(neg.shape[0])])
            l_ref, _ = T.classifier_loss_grad(th_before, x, y, d)
            assert abs(loss.item() - l_ref) < 1e-13
        _cmp(case["theta"][k], theta, 1e-9)
        _cmp(case["target"][k], target, 1e-9)


@pytest.mark.parametrize("idx", [0, 2])
def test_ratio_estimator_class_trains_like_the_reference(gold, idx):
    """RatioEstimator.optimize_callback end to end (device pool, np.random minibatch draw, kernels, module views) and the
    density ratios the trained target network then yields."""
    from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
    case = gold["ratio_training"][idx]
    d = case["d"]
    est = RatioEstimator(in_size=d, device=DEV, dtype=torch.float64, ema_weight=case["ema_weight"], lr=case["lr"],
                         weight_decay=case["weight_decay"])
    torch.nn.utils.vector_to_parameters(case["theta0"].to(DEV).clone(), est.classifier.parameters())
    est.set_neg_samples(case["neg"].to(DEV))
    est.device_permutation = False                        # parity setting: the reference's np.random minibatch draw
    np.random.seed(case["seed"])
    flat = lambda net: torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    for k, xk in enumerate(case["callbacks"]):
        est.optimize_callback(xk.clone())                 # CPU iterates are accepted (copied into the device pool)
        _cmp(case["theta"][k], flat(est.classifier), 1e-9)
        _cmp(case["target"][k], flat(est._target_network), 1e-9)
    r = est(case["x_test"].to(DEV))
    np.testing.assert_allclose(r.cpu().numpy(), case["ratio_test"].numpy(), rtol=1e-7)
    # the acquisition kernels read the flat target vector itself
    w = est.device_weights(DEV)
    assert w.W1.data_ptr() == est._train["target"].data_ptr()


def test_float32_cpu_modules_are_refreshed_after_each_step():
    from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
    torch.manual_seed(0)
    est = RatioEstimator(in_size=3, device=None, dtype=torch.float32)
    est.set_neg_samples(torch.rand(8, 3, device=DEV))
    before = [p.detach().clone() for p in est.classifier.parameters()]
    est.optimize_callback(torch.rand(16, 3, device=DEV))
    after = list(est.classifier.parameters())
    assert after[0].dtype == torch.float32 and after[0].device.type == "cpu"
    assert any(not torch.equal(a, b) for a, b in zip(after, before))
    tgt = torch.cat([p.reshape(-1) for p in est._target_network.parameters()])
    np.testing.assert_allclose(tgt.numpy(), est._train["target"].cpu().numpy(), rtol=1e-6, atol=1e-8)
    assert float(est.last_loss) > 0


def test_training_without_a_device_fails_loudly():
    from conformal_bayesopt_b200 import _lib
    from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
    est = RatioEstimator(in_size=2, dtype=torch.float64)
    est.set_neg_samples(torch.rand(4, 2, dtype=torch.float64))
    with pytest.raises(_lib.CboError):
        est.optimize_callback(torch.rand(8, 2, dtype=torch.float64))


@pytest.mark.parametrize("idx", [0, 1])
def test_sgld_update_kernel_replays_the_reference_optimizer(gold, idx):
    from conformal_bayesopt_b200 import ops
    from conformal_bayesopt_b200.optim import cosine_lr
    case = gold["sgld"][idx]
    X = case["x0"].to(DEV).clone()
    buf = torch.zeros_like(X)
    d = X.shape[-1]
    lb = torch.full((d,), -1e6, dtype=torch.float64, device=DEV)      # penalty identically zero
    ub = torch.full((d,), 1e6, dtype=torch.float64, device=DEV)
    for s in range(len(case["grads"])):
        lr = cosine_lr(s, case["lr0"], case["T_max"], case["lr0"] / 16.0)
        assert abs(lr - case["lrs"][s]) <= 1e-15 * abs(lr) + 1e-18
        ops.sgld_update(X, buf, (-case["grads"][s]).to(DEV), case["noise"][s].to(DEV), lb, ub, lr=case["lrs"][s],
                        momentum=0.1, temperature=case["scaled_temp"])
        np.testing.assert_allclose(X.cpu().numpy(), case["xs"][s].numpy(), rtol=0, atol=1e-14)
        np.testing.assert_allclose(buf.cpu().numpy(), case["bufs"][s].numpy(), rtol=0, atol=1e-14)


def test_sgld_update_box_penalty_gradient():
    from conformal_bayesopt_b200 import ops
    torch.manual_seed(5)
    X0 = torch.rand(6, 2, 3, dtype=torch.float64) * 1.2 - 0.1
    lb = torch.zeros(3, dtype=torch.float64)
    ub = torch.tensor([1.0, 0.9, 1.1], dtype=torch.float64)
    dacq = torch.randn_like(X0)
    X = X0.to(DEV).clone()
    zeros = torch.zeros_like(X)
    ops.sgld_update(X, None, dacq.to(DEV), zeros, lb.to(DEV), ub.to(DEV), lr=1.0, momentum=0.0, temperature=0.0)
    g = (X0 - X.cpu()).numpy()
    ref = -dacq.numpy() + T.bounds_penalty_grad(X0.numpy(), lb.numpy(), ub.numpy())
    np.testing.assert_allclose(g, ref, rtol=1e-9, atol=1e-12)


def _problem(N=48, d=3, seed=0):
    from conformal_bayesopt_b200.model_state import GPModelState
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(N, d, generator=g, dtype=torch.float64)
    y = torch.sin(3 * X).sum(-1) + 0.1 * torch.randn(N, generator=g, dtype=torch.float64)
    y = (y - y.mean()) / y.std()
    return GPModelState.fit_fixed(X.to(DEV), y.to(DEV), lengthscale=torch.full((d,), 0.5, dtype=torch.float64),
                                  outputscale=1.0, noise=1e-2, mean=0.0), X


def test_optimize_acqf_sgld_runs_the_reference_loop_on_the_device():
    """The default loop shape of the reference (cUCB, q = 2, ratio estimator trained by the callback, randomised)."""
    from conformal_bayesopt_b200.acquisitions import qConformalUpperConfidenceBound
    from conformal_bayesopt_b200.optim import optimize_acqf_sgld
    from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
    state, Xtr = _problem()
    d = Xtr.shape[1]
    torch.manual_seed(0)
    np.random.seed(0)
    est = RatioEstimator(in_size=d, device=DEV, dtype=torch.float64, lr=5e-3, ema_weight=2e-2, weight_decay=5e-6)
    est.set_neg_samples(Xtr.to(DEV))
    acq = qConformalUpperConfidenceBound(alpha=0.2, temp=3e-2, grid_res=32, max_grid_refinements=0, ratio_estimator=est,
                                         randomized=True, model=state, beta=0.2)
    bounds = torch.stack([torch.zeros(d), torch.ones(d)]).to(DEV, torch.float64)
    cand, val = optimize_acqf_sgld(acq, bounds, q=2, num_restarts=4, raw_samples=32, sgld_steps=48, warmup_steps=8,
                                   lr=2e-2, temperature=4e-3, options=dict(callback=est.optimize_callback, seed=1))
    assert cand.shape == (2, d) and val.shape == ()
    assert bool(((cand >= 0) & (cand <= 1)).all()) and bool(torch.isfinite(val))
    assert est._train is not None and est._train["step"] > 0          # the callback trained on the device
    assert float(est._train["target"].abs().max()) > 0
    allc, allv = optimize_acqf_sgld(acq, bounds, q=2, num_restarts=4, raw_samples=32, sgld_steps=16, warmup_steps=4,
                                    lr=2e-2, temperature=4e-3, return_best_only=False, options=dict(seed=1))
    assert allc.shape == (4, 2, d) and allv.shape == (4,)
    traj, tv = optimize_acqf_sgld(acq, bounds, q=2, num_restarts=4, raw_samples=32, sgld_steps=16, warmup_steps=4,
                                  lr=2e-2, temperature=4e-3, return_full_traj=True, options=dict(seed=1))
    assert traj.shape == (4, 13, 2, d) and tv.shape == (4, 13)


def test_sgld_improves_the_acquisition_value_without_noise():
    """temperature = 0: the loop is momentum gradient ascent on acq; the selected iterate is at least as good as the start."""
    from conformal_bayesopt_b200.acquisitions import qConformalExpectedImprovement
    from conformal_bayesopt_b200.optim import optimize_acqf_sgld
    state, Xtr = _problem(seed=3)
    d = Xtr.shape[1]
    acq = qConformalExpectedImprovement(alpha=0.2, temp=3e-2, grid_res=32, max_grid_refinements=0, ratio_estimator=None,
                                        model=state, best_f=0.5)
    torch.manual_seed(2)
    acq.fixed_probes = (2.0 * (torch.rand(state.N + 1, 8) > 0.5) - 1.0).double()
    bounds = torch.stack([torch.zeros(d), torch.ones(d)]).to(DEV, torch.float64)
    ics = (0.2 + 0.6 * torch.rand(6, 1, d, dtype=torch.float64)).to(DEV)
    with torch.no_grad():
        start = acq(ics)
    cand, val = optimize_acqf_sgld(acq, bounds, q=1, num_restarts=6, batch_initial_conditions=ics, sgld_steps=40,
                                   warmup_steps=0, lr=5e-3, temperature=0.0, return_best_only=False)
    assert bool((val >= start - 1e-12).all())
    assert float((val - start).max()) > 0


def test_graph_replay_loop_matches_the_direct_loop_bitwise():
    from conformal_bayesopt_b200.acquisitions import qConformalUpperConfidenceBound
    from conformal_bayesopt_b200.optim import optimize_acqf_sgld
    from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
    state, Xtr = _problem(seed=4)
    d = Xtr.shape[1]
    bounds = torch.stack([torch.zeros(d), torch.ones(d)]).to(DEV, torch.float64)
    ics = (0.1 + 0.8 * torch.rand(5, 2, d, dtype=torch.float64, generator=torch.Generator().manual_seed(9))).to(DEV)
    res = []
    for use_graph in (False, True):
        torch.manual_seed(7)
        np.random.seed(7)
        est = RatioEstimator(in_size=d, device=DEV, dtype=torch.float64, lr=5e-3, ema_weight=2e-2, weight_decay=5e-6)
        est.set_neg_samples(Xtr.to(DEV))
        acq = qConformalUpperConfidenceBound(alpha=0.2, temp=3e-2, grid_res=32, max_grid_refinements=0,
                                             ratio_estimator=est, randomized=True, model=state, beta=0.2)
        acq.fixed_probes = (2.0 * (torch.rand(state.N + 2, 8, generator=torch.Generator().manual_seed(3)) > 0.5) - 1.0).double()
        acq.fixed_uniforms = torch.rand(5, 32, 2, dtype=torch.float64, generator=torch.Generator().manual_seed(4)).to(DEV)
        traj, vals = optimize_acqf_sgld(acq, bounds, q=2, num_restarts=5, batch_initial_conditions=ics, sgld_steps=24,
                                        warmup_steps=4, lr=1e-2, temperature=4e-3, return_full_traj=True,
                                        options=dict(callback=est.optimize_callback, cuda_graph=use_graph,
                                                     score_batch_limit=5))    # fixed_uniforms are sized for 5 joint batches per call
        res.append((traj.clone(), vals.clone(), est._train["target"].clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])


def test_sequential_greedy_mode_uses_pending_points():
    """optimize_acqf_sgld(sequential=True): q single-point problems with the earlier picks as X_pending (optimize.py:113-149)."""
    from conformal_bayesopt_b200.acquisitions import qConformalUpperConfidenceBound
    from conformal_bayesopt_b200.optim import optimize_acqf_sgld
    state, Xtr = _problem(seed=6)
    d = Xtr.shape[1]
    acq = qConformalUpperConfidenceBound(alpha=0.2, temp=3e-2, grid_res=32, max_grid_refinements=0, ratio_estimator=None,
                                         model=state, beta=0.2)
    bounds = torch.stack([torch.zeros(d), torch.ones(d)]).to(DEV, torch.float64)
    torch.manual_seed(1)
    cand, vals = optimize_acqf_sgld(acq, bounds, q=3, num_restarts=4, raw_samples=32, sequential=True, sgld_steps=12,
                                    warmup_steps=2, lr=1e-2, temperature=4e-3, options=dict(seed=5))
    assert cand.shape == (3, d) and vals.shape == (3,)
    assert acq.X_pending is None                                   # restored
    assert bool(((cand >= 0) & (cand <= 1)).all()) and bool(torch.isfinite(vals).all())


def test_optimize_acqf_sgld_list_chains_pending_points_and_trains_the_estimator():
    from conformal_bayesopt_b200.acquisitions import qConformalExpectedImprovement, qConformalUpperConfidenceBound
    from conformal_bayesopt_b200.optim import optimize_acqf_sgld_list
    from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
    state, Xtr = _problem(seed=8)
    d = Xtr.shape[1]
    est = RatioEstimator(in_size=d, device=DEV, dtype=torch.float64, lr=5e-3, ema_weight=2e-2, weight_decay=5e-6)
    est.set_neg_samples(Xtr[:16].to(DEV))
    kw = dict(alpha=0.2, temp=3e-2, grid_res=32, max_grid_refinements=0, ratio_estimator=est, randomized=True, model=state)
    acqs = [qConformalUpperConfidenceBound(beta=0.2, **kw), qConformalExpectedImprovement(best_f=0.4, **kw)]
    bounds = torch.stack([torch.zeros(d), torch.ones(d)]).to(DEV, torch.float64)
    torch.manual_seed(2)
    np.random.seed(2)
    cand, vals = optimize_acqf_sgld_list(acqs, bounds, num_restarts=4, raw_samples=32, options=dict(seed=7), sgld_steps=16,
                                         warmup_steps=4, lr=1e-2, temperature=4e-3)
    assert cand.shape == (2, d) and vals.shape == (2,)
    assert acqs[1].X_pending is not None and torch.equal(acqs[1].X_pending, cand[:1])      # second function saw the first pick
    assert est._train is not None and est._train["step"] > 0                               # the callback was installed
    with pytest.raises(ValueError):
        optimize_acqf_sgld_list([], bounds, num_restarts=4, raw_samples=32)


@pytest.mark.parametrize("d,n_neg,n_pos", [(1, 3, 3), (2, 5, 9), (7, 64, 64), (33, 65, 1), (64, 130, 70)])
def test_train_step_kernel_edge_shapes_match_the_oracle(d, n_neg, n_pos):
    """Odd input widths (incl. the class default in_size = 1 and the kernel's maximum 64), ragged tiles, unequal class counts,
    several steps with Adam state carried over -- against oracle/training.py."""
    from conformal_bayesopt_b200 import ops
    g = torch.Generator().manual_seed(d * 1000 + n_neg)
    P = ops.ratio_num_params(d)
    th0 = torch.randn(P, generator=g, dtype=torch.float64) * 0.3
    neg = torch.rand(n_neg, d, generator=g, dtype=torch.float64)
    pool = 0.5 + 0.3 * torch.randn(2 * n_pos + 3, d, generator=g, dtype=torch.float64)
    idx = torch.randperm(pool.shape[0], generator=g)[:n_pos]
    th, m, v, tg = th0.to(DEV).clone(), torch.zeros(P, dtype=torch.float64, device=DEV), torch.zeros(P, dtype=torch.float64, device=DEV), \
        torch.zeros(P, dtype=torch.float64, device=DEV)
    rt, rm, rv, rtg = th0.numpy().copy(), np.zeros(P), np.zeros(P), np.zeros(P)
    loss = torch.zeros((), dtype=torch.float64, device=DEV)
    for step in (1, 2, 3):
        ops.ratio_train_step(th, m, v, tg, x_neg=neg.to(DEV), x_pos=pool.to(DEV), pos_idx=idx.to(DEV), n_pos=n_pos, step=step,
                             lr=5e-3, weight_decay=1e-4, ema_weight=1e-2, loss=loss)
        rt, rm, rv, rtg, rl = T.ratio_train_step(rt, rm, rv, rtg, step, neg.numpy(), pool.numpy()[idx.numpy()], d, 5e-3, 1e-4, 1e-2)
        assert abs(loss.item() - rl) < 1e-13
        np.testing.assert_allclose(th.cpu().numpy(), rt, rtol=0, atol=1e-9)
        np.testing.assert_allclose(tg.cpu().numpy(), rtg, rtol=0, atol=1e-9)
        np.testing.assert_allclose(m.cpu().numpy(), rm, rtol=1e-9, atol=1e-15)


def test_train_step_rejects_unsupported_shapes():
    from conformal_bayesopt_b200 import _lib
    lib = _lib.load()
    z = torch.zeros(8, dtype=torch.float64, device=DEV)
    st = lib.cbo_ratio_train_step_f64(65, 64, z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), 1, z.data_ptr(),
                                      None, 1, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1e-2, None, z.data_ptr(), 64, None)
    assert st == 2                                                        # CBO_ERR_UNSUPPORTED: d > 64
    st = lib.cbo_ratio_train_step_f64(2, 32, z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), 1, z.data_ptr(),
                                      None, 1, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1e-2, None, z.data_ptr(), 64, None)
    assert st == 2                                                        # hidden width other than 64
    st = lib.cbo_ratio_train_step_f64(2, 64, z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), 1, z.data_ptr(),
                                      None, 1, 0, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1e-2, None, z.data_ptr(), 64, None)
    assert st == 1                                                        # CBO_ERR_INVALID_ARG: step is 1-based


def test_sample_subset_draws_distinct_indices_on_the_device():
    """cbo_sample_subset_i64 (the device-side minibatch draw of RatioEstimator.optimize_callback): k distinct indices in
    [0, n), reproducible for a seed, different across seeds, every index about equally likely, also at the edges
    (k = n is a permutation, n = 1, n just above a power of four)."""
    from conformal_bayesopt_b200 import ops, _lib
    for n, k in [(1, 1), (2, 2), (5, 3), (257, 257), (1025, 64), (31392, 1024), (4 ** 8 + 1, 4096)]:
        a = ops.sample_subset(n, k, 1234, DEV).clone()
        b = ops.sample_subset(n, k, 1234, DEV).clone()
        c = ops.sample_subset(n, k, 1235, DEV).clone()
        assert a.shape == (k,) and torch.equal(a, b)
        assert int(a.min()) >= 0 and int(a.max()) < n and a.unique().numel() == k
        if n > 64 and k < n:
            assert not torch.equal(a, c)
    n, k, reps = 200, 50, 400
    counts = torch.zeros(n, dtype=torch.float64, device=DEV)
    for s in range(reps):
        counts.index_add_(0, ops.sample_subset(n, k, 77 + s, DEV), torch.ones(k, dtype=torch.float64, device=DEV))
    expect = reps * k / n                                  # 100 per index; binomial sd ~ 8.7
    assert float((counts - expect).abs().max()) < 6 * (expect * (1 - k / n)) ** 0.5
    chi2 = float(((counts - expect) ** 2 / (expect * (1 - k / n))).sum())
    assert 0.5 * n < chi2 < 1.6 * n
    with pytest.raises(_lib.CboError):
        ops.sample_subset(4, 5, 0, DEV)


def test_callback_minibatch_draw_on_the_device_is_seeded_and_trains():
    """Default RatioEstimator: the positive minibatch comes from cbo_sample_subset_i64 (no host RNG, no upload); two estimators
    built under the same torch seed train identically, and the loss falls on a separable problem."""
    from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
    d = 3
    g = torch.Generator().manual_seed(5)
    neg = torch.rand(64, d, dtype=torch.float64, generator=g).to(DEV)
    outs = []
    for _ in range(2):
        torch.manual_seed(11)
        est = RatioEstimator(in_size=d, device=DEV, dtype=torch.float64, lr=5e-3, ema_weight=2e-2, weight_decay=5e-6)
        assert est.device_permutation
        est.set_neg_samples(neg)
        gg = torch.Generator().manual_seed(6)
        losses = []
        for step in range(60):
            est.optimize_callback((0.2 * torch.rand(16, d, dtype=torch.float64, generator=gg) + 2.0).to(DEV))
            if est.last_loss is not None:
                losses.append(float(est.last_loss))
        outs.append((est._train["theta"].clone(), losses))
    assert torch.equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]
    assert len(outs[0][1]) > 20 and outs[0][1][-1] < 0.7 * outs[0][1][0]
