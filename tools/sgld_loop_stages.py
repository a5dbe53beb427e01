"""Stage times (graph capture / steps / trajectory scoring) of repeated optimize_acqf_sgld calls: looks for sporadic stalls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from conformal_bayesopt_b200 import GPModelState
from conformal_bayesopt_b200.acquisitions import qConformalUpperConfidenceBound
from conformal_bayesopt_b200.optim import optimize_acqf_sgld
from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
from conformal_bayesopt_b200.synthetic import make_problem
d, G, N = 8, 64, 1024
dev = torch.device("cuda:0")
p = make_problem(N, d, seed=0)
st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
torch.manual_seed(0); np.random.seed(0)
est = RatioEstimator(in_size=d, device=dev, dtype=torch.float64, lr=5e-3, ema_weight=2e-2, weight_decay=5e-6)
est.set_neg_samples(p["X"].to(dev))
acq = qConformalUpperConfidenceBound(alpha=N ** -0.35, temp=3e-2, grid_res=G, max_grid_refinements=0, ratio_estimator=est,
                                     randomized=True, model=st, beta=0.2)
acq.device_probes = os.environ.get("HOST_DRAWS") is None
est.device_permutation = os.environ.get("HOST_DRAWS") is None
bounds = torch.stack([torch.zeros(d), torch.ones(d)]).to(dev, torch.float64)
ics = torch.rand(8, 2, d, dtype=torch.float64, device=dev)
for rep in range(14):
    n = (128, 512)[rep % 2]
    tm = {}
    t0 = time.perf_counter()
    optimize_acqf_sgld(acq, bounds, q=2, num_restarts=8, batch_initial_conditions=ics, sgld_steps=n, warmup_steps=32, lr=2e-2,
                       temperature=4e-3, return_full_traj=True, options=dict(callback=est.optimize_callback, timing=tm))
    torch.cuda.synchronize()
    print(f"steps {n:4d}: total {1e3 * (time.perf_counter() - t0):8.1f} ms | setup {1e3 * tm['setup_s']:7.1f}  steps {1e3 * tm['steps_s']:7.1f}  scoring {1e3 * tm['scoring_s']:7.1f}", flush=True)
