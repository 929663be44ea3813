"""dnpy_b200 — B200-native execution backend with dnpy's Python API.

    from dnpy_b200.layers import *
    from dnpy_b200.net import *
    from dnpy_b200.optimizers import *
    from dnpy_b200 import losses, metrics, utils, initializers, regularizers

Everything numerical runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/dnpy_b200.h`` (libdnpy_b200.so, loaded with ctypes).  There is no CPU fallback.
"""
from .runtime import set_math, get_math, set_maxpool_route, get_maxpool_route, config  # noqa: F401

__version__ = "0.1.0"
