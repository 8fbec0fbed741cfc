"""Drop-in replacement for the reference's f2py module ``chimera.moduls.fimera``.

``import chimera_b200.fimera as chimera`` gives the same callables, argument order and return
conventions as the Fortran extension (reference f90/*.f90 through f2py), with
the work done by libchimera_b200.so on the GPU.  To run the unmodified reference driver on it::

    sys.modules['chimera.moduls.fimera'] = chimera_b200.fimera

Resident mode (``chimera_b200.fimera.resident(True)`` or ``CHIMERA_B200_RESIDENT=1`` in the environment): the driver's
numpy arrays live in CUDA managed memory and no call copies its arguments -- see :mod:`chimera_b200.resident`.
"""
import os
import sys

from . import _lib, resident as _resident
from .f2py_shim import build_module

_mod = build_module(_lib.load(), "chimera", "chimera_b200.fimera", adopt=_resident.adopt, adopt_input=_resident.adopt_input)
_mod.__doc__ = __doc__


def _resident_switch(on=True):
    """switch the resident mode on / off for this process (chimera_b200.resident.enable / disable)"""
    (_resident.enable if on else _resident.disable)()
    return _resident.enabled()


_mod.resident = _resident_switch
_mod.resident_enabled = _resident.enabled
_mod.ResidentArray = _resident.ResidentArray
if os.environ.get("CHIMERA_B200_RESIDENT", "") not in ("", "0"):
    _resident.enable()
_mod.device_count = _lib.device_count
_mod.kernel_launches = _lib.kernel_launches
sys.modules[__name__] = _mod
