"""CPU tests of the host-side mirror (no CUDA calls): samplers, ratio estimator, sharding, class conventions."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conformal_bayesopt_b200 import distributed as cdist
from conformal_bayesopt_b200.acquisitions import (ConformalAcquisition, qConformalExpectedImprovement,
                                                  qConformalUpperConfidenceBound)
from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
from conformal_bayesopt_b200.sampling import IIDNormalSampler


def test_sampler_reuses_base_samples_until_shape_changes():
    s = IIDNormalSampler(4, seed=3)
    a = s.draw((4, 1))
    assert torch.equal(a, s.draw((4, 1)))                 # resample=False: cached (App. B-7)
    b = s.draw((8, 1))
    assert b.shape == (8, 1)
    assert torch.equal(IIDNormalSampler(4, seed=3).draw((4, 1)), a)
    r = IIDNormalSampler(4, resample=True, seed=3)
    assert not torch.equal(r.draw((4, 1)), r.draw((4, 1)))
    s.set_base_samples(torch.ones(4, 1))
    assert torch.equal(s.draw((4, 1)), torch.ones(4, 1, dtype=torch.float64))


def test_untrained_ratio_estimator_is_one():
    from conformal_bayesopt_b200 import _lib
    r = RatioEstimator(in_size=3, dtype=torch.float64)
    assert all(float(p.abs().sum()) == 0.0 for p in r._target_network.parameters())   # ratio_estimation.py:81-86: w == 1
    r.update_target_network()
    assert any(p.abs().sum() > 0 for p in r._target_network.parameters())
    # the forward pass and the training step are kernels: CPU tensors fail loudly, there is no torch fallback
    with pytest.raises(_lib.CboError):
        r(torch.randn(5, 3, dtype=torch.float64))
    with pytest.raises(_lib.CboError):
        r.optimize_callback(torch.randn(4, 3, dtype=torch.float64))


def test_class_conventions():
    kw = dict(alpha=0.1, temp=0.03, grid_res=8, max_grid_refinements=0, ratio_estimator=None)
    ucb = qConformalUpperConfidenceBound(**kw, model=None, beta=0.2)
    assert ucb.optimistic is True and ucb.ratio_estimator is None and ucb.X_pending is None   # acquisitions.py:144
    ei = qConformalExpectedImprovement(**kw, model=None, best_f=0.5)
    assert ei.optimistic is False                                                             # acquisitions.py:81
    with pytest.raises(AttributeError):                  # acquisitions.py:42-43
        ucb(torch.zeros(2, 1, 3, dtype=torch.float64))
    with pytest.raises(TypeError):
        qConformalUpperConfidenceBound(**kw, model=None)
    with pytest.raises(NotImplementedError):
        ConformalAcquisition(**kw, model=None, objective=lambda s, X=None: s)
    ucb.set_X_pending(torch.zeros(1, 3))
    assert ucb.X_pending.shape == (1, 3)
    assert abs(ucb.beta_prime - (0.2 * torch.pi / 2) ** 0.5) < 1e-15                          # App. B-10


def test_shard_bounds_cover_the_batch():
    for b in (1, 7, 128, 1000):
        for world in (1, 2, 3, 8):
            spans = [cdist.shard_bounds(b, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _acq_standin(X):
    return torch.sin(3.0 * X.sum(dim=(-1, -2))) + 0.1 * X[..., 0, 0]


def _worker(rank, world, port, b, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        X = torch.rand(b, 1, 4, generator=g, dtype=torch.float64)
        X[5] = X[b - 2]                                   # an exact tie across shards: lowest index must win
        full = cdist.sharded_acquisition(_acq_standin, X)
        lo, hi = cdist.shard_bounds(b, rank, world)
        val, idx = cdist.sharded_argmax(_acq_standin, X[lo:hi], lo)
        shared = torch.full((3,), float(rank), dtype=torch.float64)
        cdist.broadcast_shared_inputs([shared])
        # grid-refinement statistics hook: {min, max, sum, count} of this rank's shard -> batch-global values
        stats = cdist.refine_reducer()([3.0 + rank, 40.0 - rank, 0.25 * (rank + 1), 10.0 + rank])
        results[rank] = (full, val, idx, shared, stats)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("b", [37, 64])
def test_sharded_scoring_world_size_2_gloo(b):
    world, port = 2, 29500 + (os.getpid() % 2000) + b
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, b, results), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    X = torch.rand(b, 1, 4, generator=g, dtype=torch.float64)
    X[5] = X[b - 2]
    want = _acq_standin(X)
    for rank in range(world):
        full, val, idx, shared, stats = results[rank]
        assert stats == [3.0, 40.0, 0.75, 21.0]
        assert torch.equal(full, want)                    # sharded == unsharded, bit for bit
        assert idx == int(torch.argmax(want)) and val == float(want.max())
        assert torch.equal(shared, torch.zeros(3, dtype=torch.float64))


def test_cosine_schedule_closed_form_matches_the_reference_scheduler():
    """optim.cosine_lr against the learning rates torch's CosineAnnealingLR produced around the reference's SGLD class
    (tests/golden/ref_training.pt), including the cycles past T_max."""
    import os
    from conformal_bayesopt_b200.optim import cosine_lr
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ref_training.pt"))
    for case in gold["sgld"]:
        for s, lr in enumerate(case["lrs"]):
            assert abs(cosine_lr(s, case["lr0"], case["T_max"], case["lr0"] / 16.0) - lr) <= 1e-15 * abs(lr) + 1e-18


def test_flat_parameter_views_follow_torch_parameter_order():
    from conformal_bayesopt_b200 import ops
    d = 5
    net = torch.nn.Sequential(torch.nn.Linear(d, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64), torch.nn.ReLU(),
                              torch.nn.Linear(64, 1)).double()
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    assert flat.numel() == ops.ratio_num_params(d)
    w = ops.RatioWeights.from_flat(flat, d)
    ref = ops.RatioWeights.from_sequential(net, "cpu")
    for a, b in zip((w.W1, w.b1, w.W2, w.b2, w.W3, w.b3), (ref.W1, ref.b1, ref.W2, ref.b2, ref.W3, ref.b3)):
        assert a.shape == b.shape and torch.equal(a, b)
    flat[0] = 123.0                                   # views, not copies
    assert w.W1[0, 0] == 123.0


def test_initial_conditions_keep_the_best_raw_sample():
    """gen_batch_initial_conditions (botorch's heuristic, restated): restarts inside the bounds, the best raw sample kept."""
    from conformal_bayesopt_b200.optim import gen_batch_initial_conditions
    bounds = torch.tensor([[0.0, -1.0, 2.0], [1.0, 1.0, 4.0]], dtype=torch.float64)
    target = torch.tensor([0.3, 0.2, 3.1], dtype=torch.float64)
    seen = []

    def acq(X):                                       # (n, q, d) -> (n,)
        seen.append(X)
        return -((X - target) ** 2).sum(-1).min(-1).values

    torch.manual_seed(0)
    ics = gen_batch_initial_conditions(acq, bounds, q=2, num_restarts=5, raw_samples=64, options=dict(seed=3, batch_limit=16))
    assert ics.shape == (5, 2, 3) and len(seen) == 4
    assert bool(((ics >= bounds[0]) & (ics <= bounds[1])).all())
    raw = torch.cat(seen)
    best = raw[acq(raw).argmax()]
    assert any(torch.equal(ic, best) for ic in ics)
    with pytest.raises(RuntimeError):
        gen_batch_initial_conditions(acq, bounds, q=2, num_restarts=100, raw_samples=64)


def test_sample_subset_has_no_cpu_path():
    """cbo_sample_subset_i64 (the device-side minibatch draw of RatioEstimator.optimize_callback) draws on the device only."""
    from conformal_bayesopt_b200 import _lib, ops
    with pytest.raises(_lib.CboError):
        ops.sample_subset(10, 3, 0, "cpu")
    from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
    r = RatioEstimator(in_size=3, dtype=torch.float64)
    assert r.device_permutation is True          # default; False = the reference's np.random.permutation draw (parity setting)
