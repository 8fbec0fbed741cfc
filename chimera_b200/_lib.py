"""Loader of the CUDA shared library.  There is deliberately no fallback: if the library is
missing or cannot be loaded the import fails loudly (the oracle under oracle/ is test
infrastructure and is never used by the product path)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CHIMERA_B200_LIB") or os.path.join(_HERE, "libchimera_b200.so")  # override: kernel tuning builds
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libchimera_b200.so not found at %s -- run `python -m chimera_b200.build` "
                "(nvcc, sm_100a).  chimera_b200 has no CPU fallback." % LIB_PATH
            )
        _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        _lib.chimera_last_error.restype = ctypes.c_char_p
        _lib.chimera_version.restype = ctypes.c_char_p
    return _lib


def device_count():
    n = ctypes.c_int(0)
    load().chimera_device_count(ctypes.byref(n))
    return n.value


def kernel_launches():
    n = ctypes.c_longlong(0)
    load().chimera_kernel_launches(ctypes.byref(n))
    return n.value
