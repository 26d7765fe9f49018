"""Stand-in: see oracle/shims/README.md."""
from . import filters  # noqa: F401
