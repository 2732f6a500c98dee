"""relnet_b200 -- B200-native (sm_100a) implementation of the Relation-Networks hot path.

Hand-written CUDA behind a C ABI (``include/relnet_b200.h`` -> ``librelnet_b200.so``); this package is the thin
Python host side: ``_lib`` (ctypes binding), ``ops`` (torch tensors as buffer carriers) and ``compat`` (the reference's
own call surfaces: symbol-class methods and CustomOp classes).  There is NO CPU fallback: every op raises if the CUDA
library is missing or the tensors are not on a CUDA device.
"""
from . import _lib          # noqa: F401
from . import ops           # noqa: F401
from ._lib import RelnetError   # noqa: F401

__version__ = '0.1.0'
