"""End-to-end use of the drop-in classes the way the reference's loop uses them (scripts/black_box_opt.py, reduced to its
essentials; NOT part of the hot path):

  per round:  model state from (X, y) with fixed hyper-parameters                  -> cbo_model_build_f64
              optim.optimize_acqf_sgld on the acquisition class (q = 2, restarts):
                raw candidates scored without grad, Boltzmann-selected restarts     -> cbo_conformal_acq_f64
                SGLD steps with grad (CUDA-graph replay per optimiser batch)         -> cbo_conformal_acq_grad_f64, cbo_sgld_update_f64
                ratio-estimator training by the callback after warm-up              -> cbo_ratio_train_step_f64
                re-scoring of the iterates, best in-bounds joint batch
              query the black box, append the observations

usage: python tools/bo_demo.py [rounds] [sgld_steps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from conformal_bayesopt_b200 import GPModelState
from conformal_bayesopt_b200.acquisitions import qConformalUpperConfidenceBound
from conformal_bayesopt_b200.optim import optimize_acqf_sgld
from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
from conformal_bayesopt_b200.synthetic import ackley


def run(rounds=4, sgld_steps=64, d=4, n_init=32, restarts=8, q=2, raw=2048, seed=0, verbose=True):
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


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 64)
