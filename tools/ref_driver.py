"""Run the reference's UNMODIFIED Python driver on top of a replacement `fimera` backend.

Development/test helper (build container only: /root/reference does not exist on the GPU box).
It never edits or copies reference files; it only
  * registers a synthetic package ``chimera`` whose __path__ is the reference checkout,
  * injects the compat shims the reference needs on Python 3.12 / numpy 2.x
    (inspect.getargspec, np.int, a stub h5py -- SURVEY.md section 7 step 0),
  * installs the chosen backend as ``chimera.moduls.fimera``.
"""
import inspect
import os
import sys
import types

import numpy as np

REFERENCE = os.environ.get("CHIMERA_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE, "moduls"))


def install(fimera_module):
    """Make ``from chimera.moduls.solvers import Solver`` work with `fimera_module` underneath."""
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REFERENCE)
    if not hasattr(inspect, "getargspec"):
        inspect.getargspec = inspect.getfullargspec
    if not hasattr(np, "int"):
        np.int = int
    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except Exception:
            sys.modules["h5py"] = types.ModuleType("h5py")
    for k in [k for k in sys.modules if k == "chimera" or k.startswith("chimera.")]:
        del sys.modules[k]
    pkg = types.ModuleType("chimera")
    pkg.__path__ = [REFERENCE]
    sys.modules["chimera"] = pkg
    sub = types.ModuleType("chimera.moduls")
    sub.__path__ = [os.path.join(REFERENCE, "moduls")]
    sys.modules["chimera.moduls"] = sub
    pkg.moduls = sub
    sys.modules["chimera.moduls.fimera"] = fimera_module
    sub.fimera = fimera_module
    from chimera.moduls.chimera_main import ChimeraRun
    from chimera.moduls.diagnostics import Diagnostics
    from chimera.moduls.solvers import Solver
    from chimera.moduls.species import Specie

    return types.SimpleNamespace(Solver=Solver, Specie=Specie, ChimeraRun=ChimeraRun, Diagnostics=Diagnostics)
