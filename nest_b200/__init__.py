"""nest_b200 -- B200-native implementation of NEST's per-event detector-response chain.

The product is the C-ABI shared library ``csrc/libnest_b200.so`` (include/nest_b200.h);
``capi`` is its ctypes binding and ``nestpy_like`` the host-side mirror of the reference's
runNESTvec/execNEST drivers.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
