"""B200-native full-conformal acquisition hot path of conformal-bayesopt.

Importing the package does not need a GPU; calling any op does (no CPU fallback).
"""
from . import _lib, build  # noqa: F401
from .model_state import GPModelState  # noqa: F401
from .acquisitions import (  # noqa: F401
    ConformalAcquisition,
    qConformalExpectedImprovement,
    qConformalNoisyExpectedImprovement,
    qConformalUpperConfidenceBound,
)
from .mobo_acquisitions import (  # noqa: F401
    qConformalExpectedHypervolumeImprovement,
    qConformalNoisyExpectedHypervolumeImprovement,
)
from .ratio_estimation import RatioEstimator  # noqa: F401

__all__ = ["GPModelState", "build", "_lib", "ConformalAcquisition", "qConformalUpperConfidenceBound",
           "qConformalExpectedImprovement", "qConformalNoisyExpectedImprovement",
           "qConformalExpectedHypervolumeImprovement", "qConformalNoisyExpectedHypervolumeImprovement", "RatioEstimator"]
