"""vallex_b200: B200-native drop-in for the VALL-E X `generate_audio` / `VALLE.inference` path.

Python host code (this package) over the C ABI in include/vallex_b200.h (ctypes, _capi.py) into
hand-written sm_100a CUDA (csrc/).  There is no CPU fallback: importing works anywhere, but creating
an engine requires the built library and a B200.
"""
from . import macros  # noqa: F401
from ._capi import LIB_PATH, VallexB200Error  # noqa: F401

__all__ = ["macros", "LIB_PATH", "VallexB200Error"]
