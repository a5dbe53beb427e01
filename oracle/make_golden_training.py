"""Generate tests/golden/ref_training.pt by EXECUTING the reference's own training-side code.

Runs only in the build container (needs /root/reference).  Two reference files need nothing but torch / numpy once the
`conformalbo.utils` import is satisfied by a stub (the three names it pulls in are only touched by the unused `_Dataset`
bookkeeping):
  conformalbo/ratio_estimation.py   RatioEstimator: __init__, optimize_callback (one Adam step on BCE-with-logits over
                                    [neg; pos] + EMA of the target network), update_target_network, forward
  conformalbo/optim/sgld.py         SGLD (the SGHMC update optimize_acqf_sgld applies to the candidates)
Both modules are loaded from their files UNMODIFIED with importlib; nothing is copied into this repository.
Usage:  python oracle/make_golden_training.py   (writes tests/golden/ref_training.pt)
"""
import importlib.util
import math
import os
import sys
import types
from collections import namedtuple

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference_modules():
    pkg = types.ModuleType("conformalbo")
    pkg.__path__ = []
    utils = types.ModuleType("conformalbo.utils")
    DataSplit = namedtuple("DataSplit", ["inputs", "targets"], defaults=[np.array([]), np.array([])])
    utils.DataSplit = DataSplit
    utils.update_splits = lambda train_split, val_split, test_split, new_split, holdout_ratio: (train_split, val_split, test_split)
    utils.safe_np_cat = lambda arrays, **kw: np.concatenate([a for a in arrays if len(a)], **kw)
    sys.modules["conformalbo"] = pkg
    sys.modules["conformalbo.utils"] = utils
    mods = {}
    for name, rel in (("conformalbo.ratio_estimation", "conformalbo/ratio_estimation.py"),
                      ("conformalbo.optim.sgld", "conformalbo/optim/sgld.py")):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["conformalbo.ratio_estimation"].RatioEstimator, mods["conformalbo.optim.sgld"].SGLD


def flat(net):
    return torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone()


def ratio_training_case(RatioEstimator, d, n_neg, n_steps, rows_per_callback, lr, ema, wd, seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
    est = RatioEstimator(in_size=d, device=None, dtype=torch.float64, ema_weight=ema, lr=lr, weight_decay=wd)
    neg = torch.rand(n_neg, d, dtype=torch.float64)
    est.set_neg_samples(neg)
    assert flat(est._target_network).abs().max() == 0.0       # the target network starts at zero (ratio_estimation.py:81-86)
    case = dict(d=d, seed=seed, lr=lr, ema_weight=ema, weight_decay=wd, neg=neg, theta0=flat(est.classifier),
                callbacks=[], perms=[], theta=[], target=[], trained=[])
    for s in range(n_steps):
        xk = 0.5 + 0.2 * torch.randn(rows_per_callback, d, dtype=torch.float64)
        # the positive minibatch is np.random.permutation(num_positive)[:num_negative] (ratio_estimation.py:129-131):
        # replay the generator to record the rows it will pick, then restore it so that the reference draws the same ones
        n_pos_after = len(est._pos_samples) + rows_per_callback
        st = np.random.get_state()
        perm = np.random.permutation(n_pos_after)[:n_neg]
        np.random.set_state(st)
        before = flat(est.classifier)
        est.optimize_callback(xk.clone())
        trained = not torch.equal(before, flat(est.classifier))
        if not trained:
            np.random.set_state(st)          # no step was taken: the generator was not consumed
        case["callbacks"].append(xk)
        case["perms"].append(torch.from_numpy(perm.copy()) if trained else None)
        case["trained"].append(trained)
        # full parameter vectors after the first training step and after the last callback; in between a strided sample
        keep_all = (trained and sum(case["trained"]) == 1) or s == n_steps - 1
        stride = 1 if keep_all else 37
        case["theta"].append(flat(est.classifier)[::stride].clone())
        case["target"].append(flat(est._target_network)[::stride].clone())
    x_test = torch.rand(17, d, dtype=torch.float64)
    with torch.no_grad():
        case["x_test"], case["ratio_test"] = x_test, est(x_test)
    return case


def sgld_case(SGLD, shape, lr0, temperature, n_steps, T_max, seed):
    """SGLD(momentum=.1) + CosineAnnealingLR as optimize_acqf_sgld sets them up (optim/optimize.py:199-209), driven with
    recorded gradients; the injected noise is recovered from the generator state (torch.randn_like on CPU)."""
    torch.manual_seed(seed)
    x = torch.rand(shape, dtype=torch.float64).requires_grad_(True)
    d = shape[-1]
    scaled_temp = temperature / math.sqrt(d)
    opt = SGLD([x], lr=lr0, momentum=.1, temperature=scaled_temp)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=T_max, eta_min=lr0 / 16.)
    case = dict(x0=x.detach().clone(), lr0=lr0, temperature=temperature, scaled_temp=scaled_temp, T_max=T_max,
                grads=[], noise=[], lrs=[], xs=[], bufs=[])
    for s in range(n_steps):
        g = torch.randn(shape, dtype=torch.float64)
        x.grad = g.clone()
        st = torch.get_rng_state()
        eps = torch.randn_like(g)
        torch.set_rng_state(st)
        case["lrs"].append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
        case["grads"].append(g)
        case["noise"].append(eps)
        case["xs"].append(x.detach().clone())
        case["bufs"].append(opt.state[x]["momentum_buffer"].clone())
    return case


def main():
    os.makedirs(OUT, exist_ok=True)
    RatioEstimator, SGLD = load_reference_modules()
    gold = {"ratio_training": [], "sgld": []}
    # black_box_opt.yaml:26-30 settings (lr 5e-3, ema 2e-2, wd 5e-6) and the class defaults; the first callbacks only
    # accumulate positives (num_positive < num_negative -> no step)
    gold["ratio_training"].append(ratio_training_case(RatioEstimator, 2, 24, 7, 16, 5e-3, 2e-2, 5e-6, seed=11))
    gold["ratio_training"].append(ratio_training_case(RatioEstimator, 8, 40, 6, 16, 1e-3, 1e-2, 1e-4, seed=12))
    gold["ratio_training"].append(ratio_training_case(RatioEstimator, 12, 130, 5, 150, 5e-3, 2e-2, 5e-6, seed=13))
    gold["sgld"].append(sgld_case(SGLD, (8, 2, 8), 2e-2, 4e-3, 12, 4, seed=21))
    gold["sgld"].append(sgld_case(SGLD, (3, 1, 2), 1e-3, 5e-2, 9, 64, seed=22))
    path = os.path.join(OUT, "ref_training.pt")
    torch.save(gold, path)
    print("wrote", path, os.path.getsize(path), "bytes")
    for c in gold["ratio_training"]:
        print("ratio case d=%d trained=%s |theta-theta0|=%.3e" % (c["d"], c["trained"], (c["theta"][-1] - c["theta0"]).norm()))


if __name__ == "__main__":
    main()
