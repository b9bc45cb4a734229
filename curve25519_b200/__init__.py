"""curve25519_b200 -- B200-native batched Curve25519 / Ed25519 engine.

The product is the CUDA library `libcurve25519_b200.so` and its C ABI (include/c25519_b200.h,
include/c25519_legacy.h).  This Python package is the thin host-side mirror used by tests and bench.py:

    from curve25519_b200 import api
    shared, sk_clamped = api.x25519_shared(pk, sk)          # numpy (host) or torch.cuda (device) uint8 [n,32]

PyTorch is used only for device memory, streams and torch.distributed plumbing.
"""
from . import _native  # noqa: F401
from ._native import EngineError  # noqa: F401

__all__ = ["api", "EngineError"]
