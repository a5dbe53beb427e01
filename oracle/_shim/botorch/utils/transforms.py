"""match_batch_shape / t_batch_mode_transform / concatenate_pending_points (botorch 0.7 behaviour)."""
from functools import wraps

import torch


def is_fully_bayesian(model):
    return False


def match_batch_shape(X, Y):
    return X.expand(Y.shape[:-2] + X.shape[-2:])


def t_batch_mode_transform(expected_q=None, assert_output_shape=True):
    def decorator(method):
        @wraps(method)
        def decorated(acqf, X, *args, **kwargs):
            if X.dim() < 2:
                raise ValueError(f"{type(acqf).__name__} requires X to have at least 2 dimensions, but received X with only {X.dim()} dimensions.")
            if expected_q is not None and X.shape[-2] != expected_q:
                raise AssertionError(f"Expected X to be `batch_shape x q={expected_q} x d`, but got X with shape {X.shape}.")
            X = X if X.dim() > 2 else X.unsqueeze(0)
            output = method(acqf, X, *args, **kwargs)
            if assert_output_shape and output.shape != X.shape[:-2]:
                raise AssertionError(f"Expected the output shape to match the t-batch shape {X.shape[:-2]}, got {output.shape}.")
            return output
        return decorated
    return decorator


def concatenate_pending_points(method):
    @wraps(method)
    def decorated(cls, X, **kwargs):
        if cls.X_pending is not None:
            X = torch.cat([X, match_batch_shape(cls.X_pending, X)], dim=-2)
        return method(cls, X, **kwargs)
    return decorated
