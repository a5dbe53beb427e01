"""A few device-side training steps of the SGLD loop for ncu: ratio-estimator step (k_ratio_train_grad + k_ratio_adam_ema) at
2 x 1024 minibatch rows, d = 8, and the SGHMC update (k_sgld_update) of 8 x 2 x 8 candidates."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conformal_bayesopt_b200 import ops
d, n = 8, 1024
dev = torch.device("cuda:0")
torch.manual_seed(0)
P = ops.ratio_num_params(d)
theta = torch.randn(P, dtype=torch.float64, device=dev) * 0.2
m, v, target = torch.zeros_like(theta), torch.zeros_like(theta), torch.zeros_like(theta)
neg = torch.rand(n, d, dtype=torch.float64, device=dev)
pool = 0.5 + 0.2 * torch.randn(8 * n, d, dtype=torch.float64, device=dev)
idx = torch.randperm(8 * n, device=dev)[:n].contiguous()
loss = torch.zeros((), dtype=torch.float64, device=dev)
X = torch.rand(8, 2, d, dtype=torch.float64, device=dev); buf = torch.zeros_like(X); g = torch.randn_like(X); nz = torch.randn_like(X)
lb, ub = torch.zeros(d, dtype=torch.float64, device=dev), torch.ones(d, dtype=torch.float64, device=dev)
ws = None
for step in range(1, 5):
    if step == 4:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    ws = ops.ratio_train_step(theta, m, v, target, x_neg=neg, x_pos=pool, pos_idx=idx, n_pos=n, step=step, lr=5e-3,
                              weight_decay=5e-6, ema_weight=2e-2, loss=loss, workspace=ws)
    ops.sgld_update(X, buf, g, nz, lb, ub, lr=2e-2, momentum=0.1, temperature=4e-3 / 8 ** 0.5)
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
print("ok", loss.item())
