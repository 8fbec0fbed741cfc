"""`fimera`-compatible module whose subroutines are the reference's OWN Fortran, executed by oracle/f90py.py
(TEST INFRASTRUCTURE; build container only -- it reads /root/reference/f90/*.f90).

The Python-visible interface (argument order, hidden dimensions, intent(out) allocation, return convention) is
generated from tests/golden/fimera.pyf, the signature file f2py's own front end produced from the same sources, so
neither the interface nor the arithmetic passes through the hand-written shim or the C++ oracle."""
import contextlib
import io
import os
import re
import types

import numpy as np

from .f90py import F90Module

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CHIMERA_REF", "/root/reference")
FILES = ["fb_io", "fb_math", "fb_math_env", "grid_deps", "grid_deps_env", "grid_deps_chnk", "grid_deps_env_chnk",
         "maxwell_solvers", "particle_tools", "devices", "utils", "SR"]


def _intent(v):
    it = set(v.get("intent", []))
    return "hide" if "hide" in it else ("inout" if {"in", "out"} <= it else ("out" if "out" in it else "in"))


def _dtype(v, internal=False):
    kind = str((v.get("kindselector") or {}).get("kind") or "")
    t = v["typespec"]
    if t == "complex":
        return np.dtype("complex128")
    if t == "integer":
        if internal:
            return np.dtype("int64")  # the translator keeps every integer array as int64
        return {"": np.dtype("int32"), "1": np.dtype("int8"), "4": np.dtype("int32"), "8": np.dtype("int64")}[kind]
    return np.dtype("float64")


def _ev(expr, env):
    e = re.sub(r"shape\(\s*(\w+)\s*,\s*(\d+)\s*\)", r"\1.shape[\2]", str(expr)).replace("/", "//")
    return int(eval(e, {}, env))  # noqa: S307 -- expressions come from the committed .pyf


def available():
    return os.path.isdir(os.path.join(REF, "f90"))


def load():
    from numpy.f2py import crackfortran

    with contextlib.redirect_stdout(io.StringIO()):
        blocks = crackfortran.crackfortran([os.path.join(ROOT, "tests", "golden", "fimera.pyf")])
    subs = {s["name"]: s for s in blocks[0]["body"][0]["body"]}
    f90 = F90Module([os.path.join(REF, "f90", f + ".f90") for f in FILES])
    mod = types.ModuleType("fimera_f90")
    mod.error = ValueError
    mod._f90 = f90

    def make(name, s):
        v = s["vars"]
        visible = [a for a in s["args"] if _intent(v[a]) in ("in", "inout")]
        hidden = [a for a in s["args"] if _intent(v[a]) == "hide"]

        def fn(*args):
            if len(args) != len(visible):
                raise TypeError("%s takes %d arguments (%s)" % (name, len(visible), ", ".join(visible)))
            env, orig = {}, {}
            for a, x in zip(visible, args):
                if "dimension" in v[a]:
                    want = _dtype(v[a], internal=True)
                    arr = np.asarray(x)
                    orig[a] = arr
                    if _intent(v[a]) == "inout" and arr.dtype == want and arr.flags.f_contiguous and arr.flags.writeable:
                        env[a] = arr  # f2py works in place on a conforming array
                    else:
                        env[a] = np.array(arr, dtype=want, order="F")
                else:
                    env[a] = x
            for h in hidden:
                env[h] = _ev(v[h]["="], env)
            for a in s["args"]:
                if _intent(v[a]) == "out":
                    if "dimension" in v[a]:
                        env[a] = np.zeros(tuple(_ev(d, env) for d in v[a]["dimension"]), dtype=_dtype(v[a], internal=True), order="F")
                    else:
                        env[a] = 0
            for a in visible:  # f2py's own shape checks
                if "dimension" in v[a]:
                    shape = tuple(_ev(d, env) for d in v[a]["dimension"])
                    if env[a].shape != shape:
                        raise ValueError("%s: %s has shape %r, expected %r" % (name, a, env[a].shape, shape))
            scal = getattr(f90, name)(*[env[a] for a in s["args"]])
            outs = []
            for a in s["args"]:
                it = _intent(v[a])
                if it in ("out", "inout"):
                    if "dimension" in v[a]:
                        r = env[a]
                        dt = _dtype(v[a])
                        outs.append(r if r.dtype == dt else r.astype(dt))
                    else:
                        outs.append(scal[a])
            return outs[0] if len(outs) == 1 else (tuple(outs) if outs else None)

        fn.__name__ = name
        return fn

    for name, s in subs.items():
        if name in f90.fn:
            setattr(mod, name, make(name, s))
    return mod
