"""Host wall time per forward / forward+backward through the acquisition CLASS (what a BO loop calls) against the GPU time of
the same work: shows the Python / H2D overhead around the library call (SGLD shape: 8 restarts x q=2, N=1024)."""
import math, os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import GPModelState
from conformal_bayesopt_b200.acquisitions import qConformalUpperConfidenceBound
from conformal_bayesopt_b200.ratio_estimation import RatioEstimator
from conformal_bayesopt_b200.synthetic import make_problem
N, d, G, b, q = 1024, 8, 64, 8, 2
dev = torch.device("cuda:0")
p = make_problem(N, d, seed=0)
st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
ratio = RatioEstimator(in_size=d, device=dev, dtype=torch.float64)
acqf = qConformalUpperConfidenceBound(alpha=N ** -0.35, temp=3e-2, grid_res=G, max_grid_refinements=0, ratio_estimator=ratio,
                                      randomized=True, model=st, beta=0.2)
X = torch.rand(b, q, d, dtype=torch.float64, device=dev)

def step():
    Xg = X.clone().requires_grad_(True)
    loss = -acqf(Xg).sum()
    (g,) = torch.autograd.grad(loss, Xg)
    return g

for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 200
for _ in range(n):
    step()
torch.cuda.synchronize()
print(f"class path, value+grad: {(time.perf_counter() - t0) / n * 1e3:.3f} ms per SGLD step (host wall)")
with torch.no_grad():
    for _ in range(5):
        acqf(X)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        acqf(X)
    torch.cuda.synchronize()
    print(f"class path, value only: {(time.perf_counter() - t0) / n * 1e3:.3f} ms per call (host wall)")
from conformal_bayesopt_b200.graphs import GraphedConformalAcq
from conformal_bayesopt_b200.helpers import _ratio_weights
eps = acqf.grid_sampler.draw((G, q), dev); em = acqf.sampler.draw((1, G, q), dev)
gr = GraphedConformalAcq(st, b, q, eps_grid=eps, eps_mc=em, ratio=_ratio_weights(ratio, dev), kind="ucb", alpha=N ** -0.35, temp=3e-2,
                         grid_res=G, randomized=True, optimistic=True, beta_prime=math.sqrt(0.2 * math.pi / 2))
for _ in range(5):
    gr(X)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    out = gr(X)
torch.cuda.synchronize()
print(f"CUDA-graph replay, value+grad (fresh probes and uniforms per step): {(time.perf_counter() - t0) / n * 1e3:.3f} ms per SGLD step (host wall)")
if len(sys.argv) > 1:
    pr = cProfile.Profile(); pr.enable()
    for _ in range(100):
        step()
    torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
