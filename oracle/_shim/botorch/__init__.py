"""Stand-in for the slice of botorch 0.7 the reference path touches (see ../README.md).  Test infrastructure."""
from . import settings  # noqa: F401
