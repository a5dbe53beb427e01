"""Seeded synthetic problems shared by the CPU and GPU tests."""
from conformal_bayesopt_b200.synthetic import ackley, make_inputs, make_problem as _mp  # noqa: F401
from oracle import gp as ogp


def make_problem(N, d, seed=0, **kw):
    kw.setdefault("ard_spread", 0.1)
    return _mp(N, d, seed=seed, **kw)


def oracle_state(p):
    return ogp.build_state(p["X"], p["y"], p["mean"], p["lengthscale"], p["outputscale"], p["noise"])
