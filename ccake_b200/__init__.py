"""ccake_b200: B200-native implementation of the CCAKE per-timestep SPH hot path.

The product is the CUDA shared library ``libccake_b200.so`` behind the C ABI of
``include/ccake_b200.h``; this package is the thin Python host mirror used by the
tests and ``bench.py`` (the C++ mirror is ``include/ccake_b200_host.hpp``).
"""
from .settings import DEFAULTS, derived_switches, from_yaml, to_config  # noqa: F401
from .workstation import CCAKEError, SPHWorkstation  # noqa: F401
from . import slabs  # noqa: F401,E402
