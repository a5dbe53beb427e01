"""End-to-end use of the drop-in classes the way the reference's loop uses them (scripts/black_box_opt.py +
conformalbo/optim/optimize.py:optimize_acqf_sgld, reduced to its essentials; NOT part of the hot path):

  per round:  model state from (X, y) with fixed hyper-parameters      -> cbo_model_build_f64
              raw candidates scored without grad, best restarts kept     -> cbo_conformal_acq_f64
              SGLD steps on (restarts x q x d) with grad                 -> cbo_conformal_acq_grad_f64
              re-score the iterates, pick the best joint batch, query the black box, append

usage: python tools/bo_demo.py [rounds] [sgld_steps]
"""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from conformal_bayesopt_b200 import GPModelState
from conformal_bayesopt_b200.acquisitions import qConformalUpperConfidenceBound
from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
from conformal_bayesopt_b200.synthetic import ackley


def run(rounds=4, sgld_steps=64, d=4, n_init=32, restarts=8, q=2, raw=2048, seed=0, verbose=True):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n_init, d, generator=g, dtype=torch.float64)
    f = lambda x: ackley(x) + 0.1 * torch.randn(x.shape[0], generator=g, dtype=torch.float64)
    y = f(X)
    best = [float(y.max())]
    ratio = RatioEstimator(in_size=d, device=dev, dtype=torch.float64)        # untrained: unit weights
    t_acq = 0.0
    for r in range(rounds):
        ys = (y - y.mean()) / y.std()
        state = GPModelState.fit_fixed(X, ys, 0.0, 0.3, 1.0, 1e-2, device=dev)
        acqf = qConformalUpperConfidenceBound(alpha=X.shape[0] ** -0.35, temp=3e-2, grid_res=64, max_grid_refinements=0,
                                              ratio_estimator=ratio, randomized=True, model=state, beta=0.2)
        t0 = time.perf_counter()
        with torch.no_grad():                                                 # raw samples -> initial conditions
            Xr = torch.rand(raw, q, d, generator=g, dtype=torch.float64).to(dev)
            ic = Xr[acqf(Xr).topk(restarts).indices].clone()
        lr, temperature = 2e-2, 4e-3 / math.sqrt(d)
        iterates = [ic.clone()]
        for s in range(sgld_steps):                                           # Langevin steps on -acq (optimize.py:224-235)
            ic.requires_grad_(True)
            loss = -acqf(ic).sum()
            (grad,) = torch.autograd.grad(loss, ic)
            with torch.no_grad():
                noise = torch.randn(ic.shape, generator=g, dtype=torch.float64).to(dev)
                ic = (ic - lr * grad + math.sqrt(2 * lr * temperature) * noise).clamp(0.0, 1.0)
            iterates.append(ic.clone())
        with torch.no_grad():
            cand = torch.cat(iterates)                                        # ((steps+1) restarts, q, d)
            vals = acqf(cand)
            xq = cand[vals.argmax()].cpu()
        torch.cuda.synchronize()
        t_acq += time.perf_counter() - t0
        X = torch.cat([X, xq])
        y = torch.cat([y, f(xq)])
        best.append(float(y.max()))
        if verbose:
            print(f"round {r}: N={X.shape[0]:4d}  acq* = {float(vals.max()):.4f}  best observed = {best[-1]:.4f}")
    if verbose:
        print(f"acquisition work: {t_acq:.2f} s for {rounds} rounds x ({raw} raw + {sgld_steps} SGLD steps x {restarts} restarts x q={q})")
    return best


if __name__ == "__main__":
    run(int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 64)
