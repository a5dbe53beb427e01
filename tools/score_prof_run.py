"""One forward call at the bench shape for ncu (CBO_SCORE_WS picks the score kernel).  usage: python tools/score_prof_run.py [candidates] [ratio 0|1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import PARAMS, workload_cpu
from conformal_bayesopt_b200 import GPModelState, ops
from conformal_bayesopt_b200.synthetic import make_ratio_weights

b = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
use_ratio = len(sys.argv) > 2 and sys.argv[2] == "1"
dev = torch.device("cuda:0")
prob, inp, Xq = workload_cpu(0, b)
state = GPModelState.fit_fixed(prob["X"], prob["y"], prob["mean"], prob["lengthscale"], prob["outputscale"], prob["noise"], device=dev)
ratio = ops.RatioWeights(*[t.to(dev) for t in make_ratio_weights(8)]) if use_ratio else None
kw = dict(eps_grid=inp["eps_grid"][:, 0].to(dev), z=inp["z"].to(dev), eps_mc=inp["eps_mc"][:, :, 0].to(dev), temp=3e-2,
          ratio=ratio, **PARAMS)
Xq = Xq.to(dev)
for _ in range(2):
    out = ops.conformal_acq(state, Xq, **kw)
torch.cuda.synchronize()
print("ok", float(out["acq"].sum()))
