"""Loader for the in-tree shared library (no fallback: the product is the CUDA library)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpfft_b200.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("pfft_b200: %s is missing -- run `python -m pfft_b200.build` "
                              "(there is no CPU or pure-Python fallback)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)   # RTLD_LOCAL: the test oracle loads a library with the same symbol names
    return _lib
