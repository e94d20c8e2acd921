"""msdfgen_b200 — B200-native multi-channel signed distance field generation behind msdfgen's generator API.

Layout: ``csrc/`` hand-written sm_100a kernels + the C ABI (``include/msdfgen_b200.h``); ``host/`` the C++ mirror of the
reference's ``msdfgen.h`` data model and generator signatures; ``api`` the ctypes binding used by the tests and bench.
"""
from .flat import FlatShapes  # noqa: F401
from . import synth  # noqa: F401
