"""CPU oracle for the full-conformal acquisition hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`conformal_bayesopt_b200/`) may import,
call or link anything under `oracle/`; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may, and there only as the checker / CPU arm.

Parity status: the reference's arithmetic lives in un-vendored third-party
packages (botorch==0.7, gpytorch==1.9.0, requirements.txt:3-4) that are absent from this image, so
the reference itself cannot be run.  The pure-torch reference functions (`conf_mask_to_bounds`,
`_conformal_integration`, `softplus`, `RatioEstimator.forward`) ARE executed here from
/root/reference by `oracle/make_golden.py` and pin the matching oracle functions through
`tests/golden/`.  Everything that crosses the botorch/gpytorch boundary
(`conformal_gp_regression`, `construct_conformal_bands`, `est_train_post_var`, the acquisition
classes) is "parity unpinned" in the reference (it has no tests or goldens for them); it is pinned
here by two independent restatements that must agree: `oracle/dense.py` (literal, op-by-op dense
algorithm) and `oracle/closed_form.py` (SURVEY.md App. A algebra).
"""
