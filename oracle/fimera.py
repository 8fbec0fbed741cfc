"""`fimera`-compatible module backed by the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE: the checker for the CUDA path, never the product.  Usage::

    from oracle import fimera as ofimera          # strict IEEE build
    from oracle.fimera import load; fast = load(fast=True)   # -O3 -ffast-math speed baseline
"""
import ctypes
import os
import subprocess
import sys

from chimera_b200.f2py_shim import build_module

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    """Compile liboracle.so / liboracle_fast.so from chimera_oracle.cpp (g++, OpenMP)."""
    srcs = [os.path.join(_HERE, n) for n in ("chimera_oracle.cpp", "sr_utils_oracle.cpp")]
    libs = [os.path.join(_HERE, n) for n in ("liboracle.so", "liboracle_fast.so")]
    newest = max(os.path.getmtime(s) for s in srcs)
    stale = force or any((not os.path.exists(l)) or os.path.getmtime(l) < newest for l in libs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "all"])


def load(fast=False, native=False):
    """fast: the -O3 -ffast-math build (CPU speed baseline).  native: additionally -march=native, compiled on THIS
    host (the reference Makefile:14 flags); falls back to the portable fast build when no compiler is around."""
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if fast and native:
        try:
            subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle_native.so"], stdout=subprocess.DEVNULL,
                                  stderr=subprocess.DEVNULL)
            name = "liboracle_native.so"
        except Exception:
            pass
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    lib = ctypes.CDLL(path)
    mod = build_module(lib, "oracle", "oracle_fimera_fast" if fast else "oracle_fimera")
    mod.build_name = name
    return mod


_mod = load(fast=False)
_mod.load = load
_mod.build = build
sys.modules[__name__] = _mod
