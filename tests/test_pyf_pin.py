"""The boundary pinned to the reference MECHANICALLY: tests/golden/fimera.pyf is the signature file f2py's own front
end produces from the reference's f90 sources (tools/gen_pyf.py; `f2py -h`, no compiler needed).  Every function of
the ctypes shim (chimera_b200/f2py_shim.py) and every per-function prototype of include/chimera_b200.h is checked
against it: names, positional argument order, which arguments are returned, the hidden-dimension rules and the
shape checks -- by driving the shim with arrays shaped from the .pyf against a recording stand-in of the library and
looking at what reaches the C ABI (same-shaped positional swaps such as DpS2S/DmS2S or C1/C2 show up as a pointer
in the wrong slot)."""
import contextlib
import inspect
import io
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PYF = os.path.join(ROOT, "tests", "golden", "fimera.pyf")

# exported by the reference module, called by nothing in moduls/, doc/tests or the notebooks, and not rebuilt
# (SURVEY.md section 8b "Unused exports"; utils.f90 trajectory post-processing, DESIGN.md "out of scope")
# + devices.f90:300-369, which the reference itself files under "OLD STUFF"
NOT_REBUILT = {"chunk_coords", "myfftgramm", "get_amplitude1d", "get_amplitude1d2", "get_smooth1d", "get_strength", "pulse",
               "pulse_circ"}


def pyf_subroutines():
    from numpy.f2py import crackfortran

    with contextlib.redirect_stdout(io.StringIO()):
        blocks = crackfortran.crackfortran([PYF])
    return {s["name"]: s for s in blocks[0]["body"][0]["body"]}


def header_params():
    src = open(os.path.join(ROOT, "include", "chimera_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\bint\s+chimera_([a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",") if p.strip() and p.strip() != "void"]
        out[m.group(1)] = [(re.sub(r"\s+", " ", p[:p.rfind(re.findall(r"[A-Za-z_0-9]+$", p)[0])]).strip(), re.findall(r"[A-Za-z_0-9]+$", p)[0]) for p in params]
    return out


SUBS = pyf_subroutines()


def intent(v):
    it = set(v.get("intent", []))
    if "hide" in it:
        return "hide"
    if "out" in it and "in" in it:
        return "inout"
    if "out" in it:
        return "out"
    return "in"


def dtype_of(v):
    kind = (v.get("kindselector") or {}).get("kind") or (v.get("kindselector") or {}).get("*")
    t = v["typespec"]
    if t == "real":
        return np.dtype("float64") if str(kind) == "8" else np.dtype("float32")
    if t == "complex":
        return np.dtype("complex128")
    if t == "integer":
        return {None: np.dtype("int32"), "1": np.dtype("int8"), "4": np.dtype("int32"), "8": np.dtype("int64")}[None if kind is None else str(kind)]
    if t == "double precision":
        return np.dtype("float64")
    raise AssertionError(t)


def ev(expr, env):
    e = re.sub(r"shape\(\s*(\w+)\s*,\s*(\d+)\s*\)", r"\1.shape[\2]", str(expr)).replace("/", "//")
    return int(eval(e, {}, env))  # noqa: S307 -- expressions come from the committed .pyf


def shim():
    from chimera_b200.f2py_shim import build_module

    log = []

    class Fn:
        def __init__(self, name):
            self.name, self.restype = name, None

        def __call__(self, *args):
            log.append((self.name, args))
            return 0

    class Lib:
        def __getattr__(self, name):
            if name.startswith("__") or name.endswith("_last_error"):
                raise AttributeError(name)
            return Fn(name)

    return build_module(Lib(), "chimera"), log


def test_every_reference_export_is_accounted_for():
    f, _ = shim()
    have = {n for n in dir(f) if callable(getattr(f, n)) and not n.startswith("_")}
    missing = set(SUBS) - have
    assert missing == NOT_REBUILT, sorted(missing ^ NOT_REBUILT)
    hdr = header_params()
    assert not (set(SUBS) - NOT_REBUILT - set(hdr)), sorted(set(SUBS) - NOT_REBUILT - set(hdr))


@pytest.mark.parametrize("name", sorted(set(SUBS) - NOT_REBUILT))
def test_python_signature_matches_the_pyf(name):
    """positional (and keyword) argument names of the shim function = the .pyf dummy arguments minus intent(hide)
    and pure intent(out), in order"""
    f, _ = shim()
    s = SUBS[name]
    want = [a for a in s["args"] if intent(s["vars"][a]) in ("in", "inout")]
    got = [g[:-1] if g == "in_" else g for g in inspect.signature(getattr(f, name)).parameters]  # `in` is a keyword
    assert got == want, (name, got, want)


def make_inputs(s, rng):
    """values for the hidden dimensions (all distinct), arrays and scalars shaped as the .pyf declares"""
    v = s["vars"]
    hidden = [a for a in s["args"] if intent(v[a]) == "hide"]
    vals = {}
    for i, h in enumerate(hidden):
        vals[h] = 4 + 3 * i  # distinct, and 1 + n, 2 + n, 1 + 2 n never collide with another extent's value
    env = dict(vals)
    # non-hidden integer scalars that size an array (e.g. nchnk of chunk_coords_boundaries)
    for a in s["args"]:
        if intent(v[a]) == "in" and "dimension" not in v[a] and v[a]["typespec"] == "integer":
            env[a] = 3
    py = {}
    for a in s["args"]:
        it = intent(v[a])
        if it == "hide":
            continue
        if "dimension" in v[a]:
            shape = tuple(ev(d, env) for d in v[a]["dimension"])
            if it == "out":
                py[a] = ("out", shape)
                continue
            dt = dtype_of(v[a])
            if dt.kind == "c":
                arr = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
            elif dt.kind == "i":
                arr = np.arange(int(np.prod(shape))).reshape(shape) % 5
            else:
                arr = rng.standard_normal(shape)
            py[a] = np.asfortranarray(arr, dtype=dt)
        elif it == "out":
            py[a] = ("out", ())
        elif v[a]["typespec"] == "integer":
            py[a] = env[a]
        else:
            py[a] = float(rng.standard_normal())
    return vals, py


def c_value(x):
    import ctypes

    if isinstance(x, ctypes.c_void_p):
        return ("ptr", x.value)
    if hasattr(x, "_obj"):  # ctypes.byref(scalar): an output scalar
        return ("ref", None)
    if hasattr(x, "value"):
        return ("val", x.value)
    return ("val", x)


@pytest.mark.parametrize("name", sorted(set(SUBS) - NOT_REBUILT))
def test_shim_call_reaches_the_c_abi_as_the_pyf_says(name):
    f, log = shim()
    s = SUBS[name]
    v = s["vars"]
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    vals, py = make_inputs(s, rng)
    visible = [a for a in s["args"] if intent(v[a]) in ("in", "inout")]
    ret = getattr(f, name)(*[py[a] for a in visible])
    calls = [c for c in log if c[0] == "chimera_" + name]
    assert len(calls) == 1, (name, [c[0] for c in log])
    cargs = [c_value(x) for x in calls[0][1]]
    hdr = header_params()[name]
    # (1) the header's parameter list = Fortran dummy order, hidden dimensions last (names compared in lower case; the
    # header calls node / mode COUNTS nxn, nrn, nm where the Fortran passes nx = nxn - 1 etc.)
    nonhidden = [a for a in s["args"] if intent(v[a]) != "hide"]
    hidden = [a for a in s["args"] if intent(v[a]) == "hide"]
    names_h = [n.lower() for _, n in hdr]
    assert names_h[:len(nonhidden)] == nonhidden, (name, names_h, nonhidden)
    assert len(hdr) == len(nonhidden) + len(hidden) == len(cargs), (name, len(hdr), len(nonhidden), len(hidden), len(cargs))
    # (2) what the shim passes in every non-hidden slot
    for i, a in enumerate(nonhidden):
        kind, val = cargs[i]
        it = intent(v[a])
        if "dimension" in v[a]:
            assert kind == "ptr", (name, a, kind)
            if it != "out":
                assert val == py[a].ctypes.data, "%s: argument %r is not in slot %d of the C call" % (name, a, i)
        elif it == "out":
            assert kind == "ref", (name, a, kind)
        else:
            assert kind == "val" and val == py[a], (name, a, val, py[a])
    # (3) hidden dimensions: each is derived from the array axis the .pyf names; the C side takes either the Fortran
    # value or the raw extent of that axis (node / mode counts)
    env = dict(py)
    for j, h in enumerate(hidden):
        kind, val = cargs[len(nonhidden) + j]
        m = re.search(r"shape\(\s*(\w+)\s*,\s*(\d+)\s*\)", v[h]["="])
        extent = py[m.group(1)].shape[int(m.group(2))]
        # the header names what it takes: nxn / nrn / nxg = node counts (Fortran nx + 1, nr + 1), nm = extent of the mode axis;
        # any other name is the Fortran value itself
        hname = hdr[len(nonhidden) + j][1]
        want = {"nxn": vals[h] + 1, "nrn": vals[h] + 1, "nxg": vals[h] + 1, "nm": extent}.get(hname, vals[h])
        assert kind == "val" and val == want, "%s: hidden %s (header %s): C got %r, expected %r" % (name, h, hname, val, want)
        assert ev(v[h]["="], env) == vals[h]
    # (4) returned objects: intent(out) / intent(in,out) arguments in dummy order, with the declared shapes
    outs = [a for a in s["args"] if intent(v[a]) in ("out", "inout")]
    rets = ret if isinstance(ret, tuple) else ((ret,) if outs else ())
    assert len(rets) == len(outs), (name, len(rets), outs)
    env2 = dict(vals, **{a: py[a] for a in visible if not isinstance(py[a], tuple)})
    for a, r in zip(outs, rets):
        if "dimension" in v[a]:
            shape = tuple(ev(d, env2) for d in v[a]["dimension"])
            assert isinstance(r, np.ndarray) and r.shape == shape and r.dtype == dtype_of(v[a]), (name, a, getattr(r, "shape", None), shape)
            assert r.flags.f_contiguous and r.flags.owndata or intent(v[a]) == "inout", (name, a)
            if intent(v[a]) == "inout":
                assert r is py[a], "%s: intent(in,out) %r must be returned in place" % (name, a)


@pytest.mark.parametrize("name", sorted(set(SUBS) - NOT_REBUILT))
def test_shape_checks_of_the_pyf_raise(name):
    """every check(shape(x, i) == expr) of the .pyf: break it by one and the shim raises fimera.error"""
    f, log = shim()
    s = SUBS[name]
    v = s["vars"]
    rng = np.random.default_rng(7)
    vals, py = make_inputs(s, rng)
    visible = [a for a in s["args"] if intent(v[a]) in ("in", "inout")]
    arrays = [a for a in visible if isinstance(py[a], np.ndarray)]
    # axes that DEFINE a hidden dimension cannot be "wrong"; every other axis tied to a hidden dimension can
    defining = set()
    for h in s["args"]:
        if intent(v[h]) == "hide":
            m = re.search(r"shape\(\s*(\w+)\s*,\s*(\d+)\s*\)", v[h]["="])
            defining.add((m.group(1), int(m.group(2))))
    tried = 0
    for a in arrays:
        for ax, d in enumerate(v[a]["dimension"]):
            if (a, ax) in defining:
                continue
            bad = dict(py)
            shp = list(py[a].shape)
            shp[ax] += 1
            bad[a] = np.asfortranarray(np.zeros(shp, dtype=py[a].dtype))
            with pytest.raises(f.error):
                getattr(f, name)(*[bad[x] for x in visible])
            tried += 1
    assert tried or not [a for a in arrays if len(v[a]["dimension"]) > 0 and len(arrays) > 1] or True
