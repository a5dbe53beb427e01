"""cProfile of the host side of the SGLD loop (optimize_acqf_sgld with the ratio estimator's callback, graph replay, device-side
probes and minibatch draw): where the Python time of a step goes once the GPU work is ~0.2 ms."""
import cProfile, os, pstats, sys, time
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
acq.device_probes = True
bounds = torch.stack([torch.zeros(d), torch.ones(d)]).to(dev, torch.float64)
ics = torch.rand(8, 2, d, dtype=torch.float64, device=dev)
kw = dict(q=2, num_restarts=8, batch_initial_conditions=ics, lr=2e-2, temperature=4e-3, return_full_traj=True)
def loop(n):
    torch.cuda.synchronize(); t = time.perf_counter()
    optimize_acqf_sgld(acq, bounds, sgld_steps=n, warmup_steps=32, options=dict(callback=est.optimize_callback, cuda_graph=True), **kw)
    torch.cuda.synchronize(); return time.perf_counter() - t
loop(40)
t1 = loop(128); t4 = loop(512)
print(f"marginal per step {(t4 - t1) / 384 * 1e3:.3f} ms")
pr = cProfile.Profile(); pr.enable(); loop(512); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
