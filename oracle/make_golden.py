"""Generate tests/golden/*.pt by EXECUTING the reference's own pure-torch functions.

Runs only in the build container (needs /root/reference; the GPU box does not have it).  The
reference package cannot be imported (botorch/gpytorch/hydra absent), so the four functions that
need only torch are lifted out of their source files with `ast` and exec'd unmodified:
  conformalbo/helpers.py            conf_mask_to_bounds
  conformalbo/acquisitions.py       _conformal_integration
  conformalbo/acquisition/monte_carlo.py  softplus
  conformalbo/ratio_estimation.py   RatioEstimator.forward (method body, bound to a stub object)
  tests/test_coverage.py            _create_test_case (the reference's own test-case generator)
Usage:  python oracle/make_golden.py   (writes tests/golden/ref_pure_functions.pt)
"""
import ast
import os
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def lift(path, name, cls=None):
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    body = tree.body
    if cls is not None:
        body = [n for n in body if isinstance(n, ast.ClassDef) and n.name == cls][0].body
    fn = [n for n in body if isinstance(n, ast.FunctionDef) and n.name == name][0]
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch, "np": np}
    exec(compile(mod, path, "exec"), ns)
    return ns[name]


def main():
    os.makedirs(OUT, exist_ok=True)
    c2b = lift("conformalbo/helpers.py", "conf_mask_to_bounds")
    integ = lift("conformalbo/acquisitions.py", "_conformal_integration")
    softplus = lift("conformalbo/acquisition/monte_carlo.py", "softplus")
    ratio_fwd = lift("conformalbo/ratio_estimation.py", "forward", cls="RatioEstimator")
    make_case = lift("tests/test_coverage.py", "_create_test_case")

    gold = {}
    torch.manual_seed(1234)

    # --- conf_mask_to_bounds: the reference's own test generator (hard masks), t=1 and t=2 ---
    cases = []
    for t in (1, 2):
        grid, mask, tgts, real_lb, real_ub = make_case(0.8, [64], 48, 1, t)
        lb, ub = c2b(grid, mask)
        cases.append(dict(grid=grid, mask=mask, tgts=tgts, real_lb=real_lb, real_ub=real_ub, lb=lb, ub=ub))
    # soft masks (interpolation branches), fp64, incl. an all-rejected row and edge-accepted rows
    for t in (1, 2):
        grid = torch.randn(40, 32, 1, t, dtype=torch.float64)
        centre = torch.randn(40, 1, 1, t, dtype=torch.float64) * 0.3
        width = torch.rand(40, 1, 1, t, dtype=torch.float64) * 1.5 + 0.2
        mask = torch.sigmoid((width - (grid - centre).abs()) / 0.1)
        mask[0] = 0.0           # nothing accepted -> +inf
        mask[1] = 1.0           # everything accepted -> grid edges, no interpolation
        lb, ub = c2b(grid, mask)
        cases.append(dict(grid=grid, mask=mask, lb=lb, ub=ub))
    # the 5-point example quoted in SURVEY.md §4
    grid = torch.arange(5, dtype=torch.float64).view(1, 5, 1, 1)
    mask = torch.tensor([.1, .4, .9, .6, .2], dtype=torch.float64).view(1, 5, 1, 1)
    lb, ub = c2b(grid, mask)
    cases.append(dict(grid=grid, mask=mask, lb=lb, ub=ub))
    gold["conf_mask_to_bounds"] = cases

    # --- _conformal_integration: (S,b,G,q) values, (b,G,q,t) masks ---
    cases = []
    for (S, b, G, q, t) in [(1, 5, 16, 1, 1), (3, 4, 8, 2, 1), (2, 3, 8, 2, 2)]:
        values = torch.randn(S, b, G, q, dtype=torch.float64)
        mask = torch.rand(b, G, q, t, dtype=torch.float64)
        opt = (torch.rand(b, G, q, t) > 0.4).double()
        ood = torch.rand(b, G, q, t, dtype=torch.float64) * 0.1
        logp = torch.randn(b, G, 1, 1, dtype=torch.float64) - 1.0
        alpha = 0.1
        out = integ(values, mask, logp, alpha, opt, ood)
        cases.append(dict(values=values, mask=mask, opt=opt, ood=ood, logp=logp, alpha=alpha, out=out))
    # degenerate: all-zero mask and all-zero opt (clamp_min branches)
    values = torch.randn(1, 2, 8, 1, dtype=torch.float64)
    zeros = torch.zeros(2, 8, 1, 1, dtype=torch.float64)
    logp = torch.full((2, 8, 1, 1), -30.0, dtype=torch.float64)
    out = integ(values, zeros, logp, 0.05, torch.ones_like(zeros), zeros)
    cases.append(dict(values=values, mask=zeros, opt=torch.ones_like(zeros), ood=zeros, logp=logp, alpha=0.05, out=out))
    gold["conformal_integration"] = cases

    # --- softplus ---
    x = torch.randn(7, 5, 2, dtype=torch.float64) * 3
    gold["softplus"] = [dict(x=x, dim=-1, temp=tmp, out=softplus(x, -1, tmp)) for tmp in (3e-2, 1.0, 1e-6)]

    # --- RatioEstimator.forward with an explicit MLP as the target network ---
    cases = []
    for d in (2, 8):
        torch.manual_seed(1)
        net = torch.nn.Sequential(
            torch.nn.Linear(d, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 1)
        ).double()
        with torch.no_grad():
            net[4].weight.mul_(60.0)  # spread the logits so both clamps are exercised
        stub = types.SimpleNamespace(_target_network=net)
        x = torch.randn(33, d, dtype=torch.float64) * 3
        with torch.no_grad():
            out = ratio_fwd(stub, x)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        cases.append(dict(x=x, state_dict=sd, out=out))
    gold["ratio_forward"] = cases

    path = os.path.join(OUT, "ref_pure_functions.pt")
    torch.save(gold, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
