"""The four tiny xgutils helpers the hot path touches, re-provided so that YAML configs written for
the reference resolve against this package."""
from . import sysutil, nputil, ptutil, optutil  # noqa: F401
