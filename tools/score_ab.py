"""A/B of the two q = 1 score kernels (CBO_SCORE_WS=0: warp pairs, =1: warp-specialised) on the bench shape: every
output compared, score-class time from the library's CUDA events.  usage: python tools/score_ab.py [candidates] [ratio 0|1] [randomized 0|1] [temp]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import PARAMS, workload_cpu
from conformal_bayesopt_b200 import GPModelState, _lib, ops
from conformal_bayesopt_b200.synthetic import make_ratio_weights

b = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
use_ratio = len(sys.argv) > 2 and sys.argv[2] == "1"
rand = len(sys.argv) > 3 and sys.argv[3] == "1"
temp = float(sys.argv[4]) if len(sys.argv) > 4 else 3e-2
dev = torch.device("cuda:0")
prob, inp, Xq = workload_cpu(0, b)
state = GPModelState.fit_fixed(prob["X"], prob["y"], prob["mean"], prob["lengthscale"], prob["outputscale"], prob["noise"], device=dev)
ratio = ops.RatioWeights(*[t.to(dev) for t in make_ratio_weights(8)]) if use_ratio else None
P = dict(PARAMS)
P["randomized"] = rand
G = P["grid_res"]
u = torch.rand(b, G, dtype=torch.float64, generator=torch.Generator().manual_seed(5)).to(dev) if rand else None
kw = dict(eps_grid=inp["eps_grid"][:, 0].to(dev), z=inp["z"].to(dev), eps_mc=inp["eps_mc"][:, :, 0].to(dev), temp=temp,
          ratio=ratio, u_rand=u, **P)
Xq = Xq.to(dev)
lib = _lib.load()
outs = ("target_grid", "conf_pred_mask", "q_conf_scores", "grid_logp", "ood_mask", "post_mean", "num_accepted", "post_var",
        "refit_mean", "refit_sd")
res, times = {}, {}
ms = (C.c_double * 8)(); nl = (C.c_int64 * 8)()
for mode in ("0", "1"):
    os.environ["CBO_SCORE_WS"] = mode
    res[mode] = ops.conformal_acq(state, Xq, outputs=outs, **kw)
    for _ in range(2):
        ops.conformal_acq(state, Xq, **kw)
    torch.cuda.synchronize()
    lib.cbo_timing_enable(1)
    lib.cbo_timing_collect(ms, nl)
    for _ in range(4):
        ops.conformal_acq(state, Xq, **kw)
    torch.cuda.synchronize()
    lib.cbo_timing_collect(ms, nl)
    lib.cbo_timing_enable(0)
    times[mode] = ms[4] / 4
    if mode == "1" and hasattr(lib, "cbo_debug_ws_prof"):
        pr = (C.c_ulonglong * 8)()
        lib.cbo_debug_ws_prof(pr)
        ops.conformal_acq(state, Xq, **kw)
        torch.cuda.synchronize()
        lib.cbo_debug_ws_prof(pr)
        print("ws profile (clk summed over warps): compute wait/total %.3f  helper wait/total %.3f  helper A/total %.3f  C+D/total %.3f"
              % (pr[0] / max(pr[1], 1), pr[2] / max(pr[3], 1), pr[4] / max(pr[3], 1), pr[5] / max(pr[3], 1)))
worst = 0.0
for k in res["0"]:
    a, c = res["0"][k], res["1"][k]
    same_nan = torch.equal(torch.isnan(a), torch.isnan(c))
    d = (a - c).abs()[~torch.isnan(a)]
    m = float(d.max()) if d.numel() else 0.0
    worst = max(worst, m)
    print(f"{k:16s} max|diff| {m:.3e} bit-equal {torch.equal(a[~torch.isnan(a)], c[~torch.isnan(a)])} nan-pattern-equal {same_nan}")
print(f"b={b} ratio={use_ratio} randomized={rand} temp={temp}: score ms pairs {times['0']:.3f}  ws {times['1']:.3f}  "
      f"ratio {times['1'] / times['0']:.3f}  worst diff {worst:.3e}")
if os.environ.get("CBO_AB_DETAIL"):
    a, c = res["0"]["conf_pred_mask"], res["1"]["conf_pred_mask"]
    d = (a - c).abs()
    print("per-candidate max diff (first 16):", [f"{v:.1e}" for v in d.max(dim=1).values[:16].tolist()])
    print("per-grid max diff:", [f"{v:.1e}" for v in d.max(dim=0).values.tolist()])
    print("fraction of (cand, grid) entries off by > 1e-9:", float((d > 1e-9).double().mean()))
