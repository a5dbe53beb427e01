"""Stand-in for the slice of gpytorch 1.9.0 the reference path touches (see ../README.md).  Test infrastructure."""
from .lazy import lazify, DenseLazy  # noqa: F401
from . import distributions  # noqa: F401
