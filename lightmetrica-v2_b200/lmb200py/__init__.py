"""Python driver for liblmb200.so (ctypes over the C ABI in include/lmb200.h).

Test/bench convenience only: the product is the CUDA library and the two Lightmetrica plugins;
nothing here computes anything. Importing this package never falls back to a CPU path: if the
shared library is missing, `capi.lib()` raises.
"""
from . import capi, scenes, scenedesc, distributed  # noqa: F401
