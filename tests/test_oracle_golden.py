"""Pin the oracle's pure-torch functions against outputs of the REFERENCE's own code.

tests/golden/ref_pure_functions.pt was produced by oracle/make_golden.py, which executes
conf_mask_to_bounds (helpers.py:15-91), _conformal_integration (acquisitions.py:225-248),
softplus (acquisition/monte_carlo.py:116-117), RatioEstimator.forward (ratio_estimation.py:97-99)
and the reference test generator (tests/test_coverage.py:5-35) lifted from /root/reference.
"""
import os

import torch

from oracle import pure
from oracle.gp import RatioMLP


def _gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "ref_pure_functions.pt"), weights_only=False)


def test_conf_mask_to_bounds_matches_reference(golden_dir):
    for case in _gold(golden_dir)["conf_mask_to_bounds"]:
        lb, ub = pure.conf_mask_to_bounds(case["grid"], case["mask"])
        assert lb.shape == case["lb"].shape
        # identical selection + interpolation formula; inf == inf must compare equal
        assert torch.equal(torch.isinf(lb), torch.isinf(case["lb"]))
        fin = ~torch.isinf(lb)
        assert torch.allclose(lb[fin], case["lb"][fin], rtol=1e-12, atol=0)
        assert torch.allclose(ub[fin], case["ub"][fin], rtol=1e-12, atol=0)


def test_reference_coverage_property(golden_dir):
    """tests/test_coverage.py:38-87: coverage from estimated bounds == coverage from true bounds."""
    for case in _gold(golden_dir)["conf_mask_to_bounds"][:2]:
        lb, ub = pure.conf_mask_to_bounds(case["grid"], case["mask"])
        tg = case["tgts"]
        est = ((tg >= lb) * (tg <= ub)).float().mean(dim=[0, -2])
        act = ((tg >= case["real_lb"]) * (tg <= case["real_ub"])).float().mean(dim=[0, -2])
        assert torch.allclose(act, est)


def test_known_answers():
    grid = torch.arange(5, dtype=torch.float64).view(1, 5, 1, 1)
    mask = torch.tensor([.1, .4, .9, .6, .2], dtype=torch.float64).view(1, 5, 1, 1)
    lb, ub = pure.conf_mask_to_bounds(grid, mask)
    assert abs(lb.item() - 1.2) < 1e-12 and abs(ub.item() - 3.25) < 1e-12
    lb, ub = pure.conf_mask_to_bounds(grid, torch.zeros_like(mask))
    assert lb.item() == float("inf") and ub.item() == float("inf")


def test_conformal_integration_matches_reference(golden_dir):
    for c in _gold(golden_dir)["conformal_integration"]:
        out = pure.conformal_integration(c["values"], c["mask"], c["logp"], c["alpha"], c["opt"], c["ood"])
        assert out.shape == c["out"].shape
        assert torch.allclose(out, c["out"], rtol=1e-13, atol=1e-15)


def test_softplus_matches_reference(golden_dir):
    for c in _gold(golden_dir)["softplus"]:
        assert torch.allclose(pure.softplus(c["x"], c["dim"], c["temp"]), c["out"], rtol=1e-14, atol=0)


def test_ratio_forward_matches_reference(golden_dir):
    for c in _gold(golden_dir)["ratio_forward"]:
        sd = c["state_dict"]
        mlp = RatioMLP(sd["0.weight"], sd["0.bias"], sd["2.weight"], sd["2.bias"], sd["4.weight"], sd["4.bias"])
        out = mlp(c["x"])
        assert torch.allclose(out, c["out"], rtol=1e-12, atol=0)
        assert out.max() > 1e5 and out.min() < 1e-5  # both clamps exercised
