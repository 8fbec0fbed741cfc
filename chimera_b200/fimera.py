"""Drop-in replacement for the reference's f2py module ``chimera.moduls.fimera``.

``import chimera_b200.fimera as chimera`` gives the same callables, argument order and return
conventions as the Fortran extension (reference f90/*.f90 through f2py), with
the work done by libchimera_b200.so on the GPU.  To run the unmodified reference driver on it::

    sys.modules['chimera.moduls.fimera'] = chimera_b200.fimera
"""
import sys

from . import _lib
from .f2py_shim import build_module

_mod = build_module(_lib.load(), "chimera", "chimera_b200.fimera")
_mod.__doc__ = __doc__
_mod.device_count = _lib.device_count
_mod.kernel_launches = _lib.kernel_launches
sys.modules[__name__] = _mod
