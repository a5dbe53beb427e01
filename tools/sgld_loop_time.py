"""Wall time of the reference's default inner loop run on the device (`optim.optimize_acqf_sgld`: cUCB, 8 restarts x q = 2,
ratio estimator trained by the callback after warm-up, randomised; conformalbo/optim/optimize.py:196-235,
hydra_config/black_box_opt.yaml).  Prints ms per SGLD step and its parts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from conformal_bayesopt_b200 import GPModelState, ops
from conformal_bayesopt_b200.acquisitions import qConformalUpperConfidenceBound
from conformal_bayesopt_b200.optim import optimize_acqf_sgld
from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
from conformal_bayesopt_b200.synthetic import make_problem

d, G = 8, 64
dev = torch.device("cuda:0")
steps = int(os.environ.get("SGLD_STEPS", "128"))
for N in (80, 1024):
    p = make_problem(N, d, seed=0)
    st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
    torch.manual_seed(0); np.random.seed(0)
    est = RatioEstimator(in_size=d, device=dev, dtype=torch.float64, lr=5e-3, ema_weight=2e-2, weight_decay=5e-6)
    est.set_neg_samples(p["X"].to(dev))
    est.device_permutation = False      # first the reference's own host draw (np.random.permutation), then the device draw (the default)
    acq = qConformalUpperConfidenceBound(alpha=N ** -0.35, temp=3e-2, grid_res=G, max_grid_refinements=0,
                                         ratio_estimator=est, randomized=True, model=st, beta=0.2)
    bounds = torch.stack([torch.zeros(d), torch.ones(d)]).to(dev, torch.float64)
    ics = torch.rand(8, 2, d, dtype=torch.float64, device=dev)
    kw = dict(q=2, num_restarts=8, batch_initial_conditions=ics, lr=2e-2, temperature=4e-3, return_full_traj=True)
    def loop(n, graph):
        torch.cuda.synchronize(); t = time.perf_counter()
        optimize_acqf_sgld(acq, bounds, sgld_steps=n, warmup_steps=32, options=dict(callback=est.optimize_callback, cuda_graph=graph), **kw)
        torch.cuda.synchronize(); return time.perf_counter() - t
    loop(40, True); loop(40, False)
    # marginal cost per step: (t(4 s) - t(s)) / (3 s) removes the set-up (graph capture) and leaves the final trajectory scoring in
    t_all = loop(steps, True)
    slope_g = (loop(4 * steps, True) - t_all) / (3 * steps)
    t_nog = loop(steps, False)
    slope_n = (loop(4 * steps, False) - t_nog) / (3 * steps)
    est.device_permutation, acq.device_probes = True, True      # no host RNG in the step
    loop(40, True)
    t_dev = loop(steps, True)
    slope_dev = (loop(4 * steps, True) - t_dev) / (3 * steps)
    est.device_permutation, acq.device_probes = False, False
    # parts, each timed alone over 100 repetitions
    def timed(fn, n=100):
        for _ in range(5): fn()
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3
    X = ics.clone(); buf = torch.zeros_like(X); noise = torch.randn_like(X)
    t_cb = timed(lambda: est.optimize_callback(X))
    est.device_permutation = True
    t_cbd = timed(lambda: est.optimize_callback(X))
    est.device_permutation = False
    t_vg = timed(lambda: acq.value_and_grad(X))
    _, g = acq.value_and_grad(X)
    t_up = timed(lambda: (torch.randn_like(X), ops.sgld_update(X, buf, g, noise, bounds[0].contiguous(), bounds[1].contiguous(), lr=1e-9, momentum=0.1, temperature=0.0)))
    print(f"N={N:5d}: {steps} SGLD steps incl. final trajectory scoring {t_all*1e3:.1f} ms = {t_all/steps*1e3:.3f} ms/step | "
          f"ratio training callback {t_cb:.3f} ms ({t_cbd:.3f} ms with device_permutation, pool of {est.num_positive} rows), value+grad {t_vg:.3f} ms, noise+update {t_up:.3f} ms; "
          f"without graph replay {t_nog/steps*1e3:.3f} ms/step; marginal per step {slope_g*1e3:.3f} ms (graph) / {slope_n*1e3:.3f} ms (direct) / {slope_dev*1e3:.3f} ms (graph, device-side probes and minibatch draw); trained steps {est._train['step']}, loss {float(est.last_loss):.4f}")
