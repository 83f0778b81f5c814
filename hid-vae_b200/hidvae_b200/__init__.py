"""hidvae_b200 -- B200-native residual-quantisation hot path of HiD-VAE (sm_100a CUDA behind a C ABI).

Importing this package loads libhidvae_b200.so; it raises ImportError when the library is not built.
"""
from . import _lib, ops  # noqa: F401
from ._lib import HidvaeError, device_info, version  # noqa: F401

__all__ = ["ops", "HidvaeError", "device_info", "version"]
