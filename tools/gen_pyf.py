#!/usr/bin/env python
"""Regenerate tests/golden/fimera.pyf: the machine-readable signature of the reference's f2py module, produced by
f2py's own front end (crackfortran needs no Fortran compiler) from the reference's sources where they lie.

    python tools/gen_pyf.py            (build container only: needs /root/reference)

tests/test_pyf_pin.py checks chimera_b200/f2py_shim.py and include/chimera_b200.h against this file."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CHIMERA_REF", "/root/reference")
# the module list of the reference Makefile (f90/*.f90 -> fimera), SURVEY.md section 8b
FILES = ["fb_io", "fb_math", "fb_math_env", "grid_deps", "grid_deps_env", "grid_deps_chnk", "grid_deps_env_chnk",
         "maxwell_solvers", "particle_tools", "devices", "utils", "SR"]


def main():
    out = os.path.join(ROOT, "tests", "golden", "fimera.pyf")
    with tempfile.TemporaryDirectory() as tmp:
        pyf = os.path.join(tmp, "fimera.pyf")
        subprocess.check_call([sys.executable, "-m", "numpy.f2py", "-h", pyf, "-m", "fimera", "--overwrite-signature"]
                              + [os.path.join(REF, "f90", f + ".f90") for f in FILES], stdout=subprocess.DEVNULL)
        text = open(pyf).read().replace(REF + "/", "")
    open(out, "w").write(text)
    print("wrote", out, "(%d lines)" % text.count("\n"))


if __name__ == "__main__":
    main()
