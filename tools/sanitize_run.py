"""Small end-to-end invocations of every kernel for compute-sanitizer (memcheck / racecheck)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import GPModelState, ops
from conformal_bayesopt_b200.synthetic import make_inputs, make_problem, make_ratio_weights
dev = torch.device("cuda:0")
for (N, d, b, q, G, temp, refine) in [(150, 4, 37, 1, 24, 3e-2, 1), (150, 4, 9, 1, 24, 0.0, 0), (70, 3, 6, 2, 10, 3e-2, 1), (40, 2, 5, 3, 8, 0.0, 0)]:
    p = make_problem(N, d, seed=1)
    inp = make_inputs(N, d, b, q, G, S=2, seed=1)
    st = GPModelState.fit_fixed(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"], device=dev)
    rw = ops.RatioWeights(*[t.to(dev) for t in make_ratio_weights(d)])
    out = ops.conformal_acq(st, inp["Xq"].to(dev), eps_grid=inp["eps_grid"].to(dev), z=inp["z"].to(dev), eps_mc=inp["eps_mc"].to(dev),
                            u_rand=inp["u_rand"].to(dev), ratio=rw, outputs=("target_grid", "conf_pred_mask", "grid_logp", "ood_mask", "post_mean", "post_var", "num_accepted"),
                            kind="ucb", alpha=0.1, temp=temp, grid_res=G, max_grid_refinements=refine, randomized=True, optimistic=True, beta_prime=1.0)
    if q == 1:
        lb, ub = ops.conf_mask_to_bounds(out["target_grid"][:, :, None, None], out["conf_pred_mask"][:, :, None, None])
    torch.cuda.synchronize()
    print(N, q, float(out["acq"].sum()))
print("sanitize_run ok")
