"""End-to-end use of the drop-in classes the way the reference's loop uses them (scripts/black_box_opt.py, reduced to its
essentials; NOT part of the hot path):

  per round:  model state from (X, y) with fixed hyper-parameters                  -> cbo_model_build_f64
              optim.optimize_acqf_sgld on the acquisition class (q = 2, restarts):
                raw candidates scored without grad, Boltzmann-selected restarts     -> cbo_conformal_acq_f64
                SGLD steps with grad (CUDA-graph replay per optimiser batch)         -> cbo_conformal_acq_grad_f64, cbo_sgld_update_f64
                ratio-estimator training by the callback after warm-up              -> cbo_ratio_train_step_f64
                re-scoring of the iterates, best in-bounds joint batch
              query the black box, append the observations

usage: python tools/bo_demo.py [rounds] [sgld_steps] [acq_fn: cucb | cei | cnei | cehvi | cnehvi | all]

Every acquisition of the family goes through the same optimiser: the reference routes any acquisition that carries a ratio
estimator to optimize_acqf_sgld (conformalbo/utils.py:243-246), so all of them need d acq / d X.
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from conformal_bayesopt_b200 import GPModelState
from conformal_bayesopt_b200.acquisitions import (qConformalExpectedImprovement, qConformalNoisyExpectedImprovement,
                                                  qConformalUpperConfidenceBound)
from conformal_bayesopt_b200.mobo_acquisitions import (box_decomposition_2d, qConformalExpectedHypervolumeImprovement,
                                                       qConformalNoisyExpectedHypervolumeImprovement)
from conformal_bayesopt_b200.optim import optimize_acqf_sgld
from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
from conformal_bayesopt_b200.synthetic import ackley


def _second_objective(x):
    """A smooth competing objective for the multi-objective demos (maximised away from the Ackley optimum)."""
    return -((x - 0.8) ** 2).sum(-1) * 4.0


def make_acqf(name, states, ratio, X, ys, dev, seed):
    """The acquisition object of one BO round, constructed the way conformalbo/utils.py:430-447 does."""
    common = dict(alpha=X.shape[0] ** -0.35, temp=3e-2, grid_res=64, max_grid_refinements=0, ratio_estimator=ratio,
                  randomized=True)
    if name == "cucb":
        return qConformalUpperConfidenceBound(model=states[0], beta=0.2, **common)
    if name == "cei":
        return qConformalExpectedImprovement(model=states[0], best_f=float(ys[:, 0].max()), **common)
    if name == "cnei":
        return qConformalNoisyExpectedImprovement(model=states[0], X_baseline=X.to(dev), prune_baseline=True, prune_seed=seed,
                                                  **common)
    ref = torch.tensor([-3.0, -3.0], dtype=torch.float64)
    if name == "cehvi":
        lo, up = box_decomposition_2d(ref, ys)                 # scripts/bbo_multi_obj.py:240-250 (there: the posterior mean)
        return qConformalExpectedHypervolumeImprovement(model=states, ref_point=ref, partitioning=(lo, up), **common)
    if name == "cnehvi":
        return qConformalNoisyExpectedHypervolumeImprovement(model=states, ref_point=ref, X_baseline=X.to(dev),
                                                             prune_baseline=True, prune_seed=seed, **common)
    raise ValueError(name)


def run(rounds=4, sgld_steps=64, d=4, n_init=32, restarts=8, q=2, raw=2048, seed=0, verbose=True, acq="cucb"):
    if acq != "cucb":
        return run_family(acq, rounds, sgld_steps, d, n_init, restarts, q, raw, seed, verbose)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    X = torch.rand(n_init, d, generator=g, dtype=torch.float64)
    f = lambda x: ackley(x) + 0.1 * torch.randn(x.shape[0], generator=g, dtype=torch.float64)
    y = f(X)
    best = [float(y.max())]
    bounds = torch.stack([torch.zeros(d), torch.ones(d)]).to(dev, torch.float64)
    t_acq = 0.0
    for r in range(rounds):
        ys = (y - y.mean()) / y.std()
        state = GPModelState.fit_fixed(X, ys, 0.0, 0.3, 1.0, 1e-2, device=dev)
        ratio = RatioEstimator(in_size=d, device=dev, dtype=torch.float64, lr=5e-3, ema_weight=2e-2, weight_decay=5e-6)
        ratio.set_neg_samples(X.to(dev))                                     # scripts/black_box_opt.py: the training inputs
        acqf = qConformalUpperConfidenceBound(alpha=X.shape[0] ** -0.35, temp=3e-2, grid_res=64, max_grid_refinements=0,
                                              ratio_estimator=ratio, randomized=True, model=state, beta=0.2)
        t0 = time.perf_counter()
        xq, val = optimize_acqf_sgld(acqf, bounds, q=q, num_restarts=restarts, raw_samples=raw, sgld_steps=sgld_steps,
                                     warmup_steps=min(32, sgld_steps // 4), lr=2e-2, temperature=4e-3,
                                     options=dict(callback=ratio.optimize_callback, seed=seed + r))
        torch.cuda.synchronize()
        t_acq += time.perf_counter() - t0
        xq = xq.cpu()
        X = torch.cat([X, xq])
        y = torch.cat([y, f(xq)])
        best.append(float(y.max()))
        if verbose:
            trained = ratio._train["step"] if ratio._train is not None else 0
            print(f"round {r}: N={X.shape[0]:4d}  acq* = {float(val):.4f}  best observed = {best[-1]:.4f}  "
                  f"(ratio estimator: {trained} training steps)")
    if verbose:
        print(f"acquisition work: {t_acq:.2f} s for {rounds} rounds x ({raw} raw + {sgld_steps} SGLD steps x {restarts} restarts x q={q})")
    return best


def run_family(acq, rounds=2, sgld_steps=32, d=4, n_init=32, restarts=4, q=2, raw=256, seed=0, verbose=True):
    """The same loop for any member of the family; two objectives for cEHVI / cNEHVI.  Returns the per-round acquisition
    optimum and the SGLD trajectory check (the selected batch is never worse than the initial conditions' best)."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    np.random.seed(seed)
    multi = acq in ("cehvi", "cnehvi")
    X = torch.rand(n_init, d, generator=g, dtype=torch.float64)
    def f(x):
        cols = [ackley(x) + 0.1 * torch.randn(x.shape[0], generator=g, dtype=torch.float64)]
        if multi:
            cols.append(_second_objective(x) + 0.05 * torch.randn(x.shape[0], generator=g, dtype=torch.float64))
        return torch.stack(cols, -1)
    Y = f(X)
    bounds = torch.stack([torch.zeros(d), torch.ones(d)]).to(dev, torch.float64)
    out = []
    for r in range(rounds):
        ys = (Y - Y.mean(0)) / Y.std(0)
        states = [GPModelState.fit_fixed(X, ys[:, k], 0.0, 0.3, 1.0, 1e-2, device=dev) for k in range(Y.shape[1])]
        ratio = RatioEstimator(in_size=d, device=dev, dtype=torch.float64, lr=5e-3, ema_weight=2e-2, weight_decay=5e-6)
        ratio.set_neg_samples(X.to(dev))
        acqf = make_acqf(acq, states, ratio, X, ys, dev, seed + r)
        t0 = time.perf_counter()
        xq, val = optimize_acqf_sgld(acqf, bounds, q=q, num_restarts=restarts, raw_samples=raw, sgld_steps=sgld_steps,
                                     warmup_steps=min(32, sgld_steps // 4), lr=2e-2, temperature=4e-3,
                                     options=dict(callback=ratio.optimize_callback, seed=seed + r))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        xq = xq.cpu()
        assert xq.shape == (q, d) and torch.isfinite(xq).all() and bool(((xq >= 0) & (xq <= 1)).all())
        X = torch.cat([X, xq])
        Y = torch.cat([Y, f(xq)])
        out.append(float(val))
        if verbose:
            print(f"[{acq}] round {r}: N={X.shape[0]:4d}  acq* = {float(val):.5f}  ({dt * 1e3:.0f} ms for {raw} raw + "
                  f"{sgld_steps} SGLD steps x {restarts} restarts x q={q})")
    return out


if __name__ == "__main__":
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    which = sys.argv[3] if len(sys.argv) > 3 else "cucb"
    for name in (["cucb", "cei", "cnei", "cehvi", "cnehvi"] if which == "all" else [which]):
        if name == "cucb" and which != "all":
            run(rounds, steps)
        else:
            run_family(name, rounds, steps)
