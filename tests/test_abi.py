"""CPU: the C-ABI boundary.  libchimera_b200.so loads without a GPU, exports every symbol that
include/chimera_b200.h declares, and its compute entry points fail loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "chimera_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return sorted(set(re.findall(r"\b(chimera_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_fimera_hot_path():
    names = declared_symbols()
    # one entry point per Fortran subroutine the reference driver calls on the path
    for n in ("push_velocs push_coords genparts sortpartsout chunk_coords_boundaries align_data_vec align_data_scl "
              "dep_curr dep_dens proj_fld eb_correction dep_curr_chnk dep_dens_chnk dep_curr_env dep_dens_env proj_fld_env "
              "eb_correction_env dep_curr_env_chnk dep_dens_env_chnk fb_vec_in fb_scl_in fb_vec_out fb_scl_out fb_eb_out "
              "fb_filtr fb_rot fb_grad fb_div fb_graddiv fb_rot_env fb_grad_env fb_div_env fb_graddiv_env "
              "maxwell_push_with_spchrg maxwell_push_wo_spchrg maxwell_init_push poiss_corr poiss_corr_stat field_drift "
              "omp_mult_vec omp_mult_scl omp_add_vec omp_add_scl undul_analytic undul_analytic_taper undul_mapped "
              "undul_mapped_tap planewave gaussbeam").split():
        assert "chimera_" + n in names, n


def test_library_exports_every_declared_symbol():
    from chimera_b200 import _lib

    lib = _lib.load()
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.chimera_version().decode()


def test_shim_covers_every_per_function_entry_point():
    import chimera_b200.fimera as f

    skip = ("last_error", "version", "device_count", "set_device", "sync", "kernel_launches", "host_traffic", "bench_gemm",
            "gemm_profile", "gemm_profile_read", "fused_profile", "fused_profile_read", "host_register", "host_unregister",
            # resident mode (chimera_b200/resident.py binds these, they are not fimera functions)
            "managed_alloc", "managed_realloc", "managed_free", "managed_owns", "managed_trim", "managed_touched", "is_device_accessible", "fill", "copy",
            "add_inplace")
    for sym in declared_symbols():
        name = sym[len("chimera_"):]
        if name.startswith("engine_") or name in skip:
            continue
        assert hasattr(f, name), name


def test_compute_fails_loudly_without_a_gpu():
    import chimera_b200.fimera as f

    if f.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(f.error):
        f.push_velocs(np.zeros((3, 4), order="F"), np.zeros((6, 4), order="F"), 0.1)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "chimera_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "liboracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, fn


def header_arity():
    """number of parameters of every `int chimera_<name>(...)` prototype in the header"""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\bint\s+(chimera_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        params = m.group(2).strip()
        out[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def test_shim_passes_as_many_arguments_as_the_header_declares():
    """static guard against drift between include/chimera_b200.h and the ctypes call sites of f2py_shim.py (ctypes
    does not check arity): every `call("<name>", ...)` with a literal name and no starred argument is compared"""
    import ast

    arity = header_arity()
    src = open(os.path.join(ROOT, "chimera_b200", "f2py_shim.py")).read()
    checked = 0
    for node in ast.walk(ast.parse(src)):
        if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id == "call" and node.args):
            continue
        first = node.args[0]
        if not (isinstance(first, ast.Constant) and isinstance(first.value, str)):
            continue
        if any(isinstance(a, ast.Starred) for a in node.args):
            continue
        sym = "chimera_" + first.value
        assert sym in arity, sym
        assert len(node.args) - 1 == arity[sym], (sym, len(node.args) - 1, arity[sym])
        checked += 1
    assert checked >= 15, checked


def test_generated_wrappers_pass_as_many_arguments_as_the_header_declares():
    """the shim functions produced by factories (omp_*, devices, SR, spectral calculus, deposits ...) against a
    recording stand-in for the library: argument counts as declared in include/chimera_b200.h"""
    from chimera_b200.f2py_shim import build_module
    from util import crandn, particles, setup

    seen = {}

    class Fn:
        def __init__(self, name):
            self.name, self.restype = name, None

        def __call__(self, *args):
            seen[self.name] = len(args)
            return 0

    class Lib:
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return Fn(name)

    f = build_module(Lib(), "chimera")
    rng = np.random.default_rng(0)
    for name in ("real_m2", "env_m3"):
        S = setup(name)
        a = S.Args
        env = "_env" if S.env else ""
        x, p, w = particles(S, 50, 1, inside_only=True)
        ind = np.array([0, 10, 20, 30, 50], dtype=np.int32)
        v, sc = crandn(rng, S.shape_fb + (3,)), crandn(rng, S.shape_fb)
        Dp, Dm, kx = a["FBDiff"]
        getattr(f, "dep_curr" + env)(x, p, w, S.zeros_sp(3), a["leftX"], *a["DepProj"])
        getattr(f, "dep_dens" + env)(x, w, S.zeros_sp(), a["leftX"], *a["DepProj"])
        getattr(f, "dep_curr" + env + "_chnk")(x, p, w, S.zeros_sp(3), ind, 3, a["leftX"], *a["DepProj"])
        getattr(f, "dep_dens" + env + "_chnk")(x, w, S.zeros_sp(), ind, 3, a["leftX"], *a["DepProj"])
        getattr(f, "proj_fld" + env)(x, w, S.zeros_sp(6), np.zeros((6, 50), order="F"), a["leftX"], *a["DepProj"])
        getattr(f, "eb_correction" + env)(S.zeros_sp(6))
        getattr(f, "fb_rot" + env)(S.zeros_fb(3), v, Dp, Dm, kx)
        getattr(f, "fb_grad" + env)(S.zeros_fb(3), sc, Dp, Dm, kx)
        getattr(f, "fb_div" + env)(S.zeros_fb(), v, Dp, Dm, kx)
        getattr(f, "fb_graddiv" + env)(v, Dp, Dm, kx)
    S = setup("real_m2")
    a = S.Args
    v, sc = crandn(rng, S.shape_fb + (3,)), crandn(rng, S.shape_fb)
    f.fb_vec_in(S.zeros_fb(3), S.zeros_sp(3), a["leftX"], *a["FBIn"])
    f.fb_scl_in(S.zeros_fb(), S.zeros_sp(), a["leftX"], *a["FBIn"])
    f.fb_vec_out(v, a["leftX"], *a["FBout"])
    f.fb_scl_out(sc, a["leftX"], *a["FBout"])
    f.fb_eb_out(S.zeros_sp(6), crandn(rng, S.shape_fb + (6,)), v, a["leftX"], *a["FBout"])
    f.omp_mult_vec(v, np.ones(S.shape_fb)); f.omp_mult_scl(sc, np.ones(S.shape_fb))
    f.omp_add_vec(v, v.copy(order="F")); f.omp_add_scl(sc, sc.copy(order="F"))
    x3, fld = np.zeros((3, 4), order="F"), np.zeros((6, 4), order="F")
    f.undul_analytic(x3, fld, 0.1, np.ones(4)); f.undul_analytic_taper(x3, fld, 0.1, np.ones(5))
    f.undul_mapped(x3, fld, 0.1, np.ones((2, 5), order="F"), np.ones(3))
    f.undul_mapped_tap(x3, fld, 0.1, np.ones((2, 5), order="F"), np.ones(5))
    f.planewave(x3, fld, 0.1, np.ones(7)); f.gaussbeam(x3, fld, 0.1, 1.0, np.ones(8))
    tr, om, g2, g3 = np.zeros((3, 6, 2), order="F"), np.ones(4), np.ones(2), np.ones(3)
    sp = np.zeros((4, 2, 3), order="F")
    f.sr_calc_far_tot(sp, tr, tr, tr, np.ones(2), 0.1, om, g2, g2, g3, g3)
    f.sr_calc_far_comp(sp, tr, tr, tr, np.ones(2), 1, 0.1, om, g2, g2, g3, g3)
    f.sr_calc_near_tot(sp, tr, tr, np.ones(2), 0.1, om, g2, g3, 5.0)
    f.sr_calc_near_comp(sp, tr, tr, np.ones(2), 2, 0.1, om, g2, g3, 5.0)
    f.sr_calc_nearcirc_tot(sp, tr, tr, np.ones(2), 0.1, om, g2, g3, g3, 5.0)
    f.sr_calc_nearcirc_comp(sp, tr, tr, np.ones(2), 3, 0.1, om, g2, g3, g3, 5.0)
    arity = header_arity()
    assert len(seen) >= 40, sorted(seen)
    for sym, n in seen.items():
        assert n == arity[sym], (sym, n, arity[sym])


def test_engine_wrapper_passes_as_many_arguments_as_the_header_declares():
    """the same static guard for chimera_b200/engine.py (self.lib.chimera_engine_*(...))"""
    import ast

    arity = header_arity()
    src = open(os.path.join(ROOT, "chimera_b200", "engine.py")).read()
    checked = 0
    for node in ast.walk(ast.parse(src)):
        if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute) and node.func.attr.startswith("chimera_")):
            continue
        if any(isinstance(a, ast.Starred) for a in node.args):
            continue
        sym = node.func.attr
        if sym in ("chimera_last_error", "chimera_version"):  # const char* f(void): not an `int` prototype
            continue
        assert sym in arity, sym
        assert len(node.args) == arity[sym], (sym, len(node.args), arity[sym])
        checked += 1
    assert checked >= 25, checked


def test_engine_config_struct_matches_the_header():
    """chimera_engine_config (header) and EngineConfig (ctypes) field by field: name, order and C type; the phase
    enum against the PHASES tuple"""
    import ctypes

    from chimera_b200.engine import PHASES, EngineConfig

    src = open(HEADER).read()
    body = re.search(r"typedef struct chimera_engine_config \{(.*?)\} chimera_engine_config;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", " ", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        ctype, names = decl.split(None, 1)
        fields += [(n.strip(), ctype) for n in names.split(",")]
    cmap = {"int": ctypes.c_int, "double": ctypes.c_double, "chb_i64": ctypes.c_longlong}
    got = [(n, t) for n, t in EngineConfig._fields_]
    assert [n for n, _ in got] == [n for n, _ in fields], ([n for n, _ in got], [n for n, _ in fields])
    for (n, t), (_, ct) in zip(got, fields):
        assert t is cmap[ct], (n, t, ct)
    enum = re.search(r"enum chimera_engine_phase \{(.*?)\};", src, flags=re.S).group(1)
    enum = re.sub(r"/\*.*?\*/", " ", enum, flags=re.S)
    ids = {m.group(1): int(m.group(2)) for m in re.finditer(r"CHB_([A-Z_]+)\s*=\s*(\d+)", enum)}
    assert ids.pop("NPHASES") == len(PHASES)
    names = {"PUSH_COORDS": "push_coords", "SORT": "sort", "DEPOSIT_J": "deposit_J", "DEPOSIT_RHO": "deposit_rho",
             "DEPOSIT_BG": "deposit_bg", "FB_IN_J": "fb_in_J", "FB_IN_RHO": "fb_in_rho", "POISSON": "poisson",
             "MAXWELL": "maxwell", "INIT_PUSH": "init_push", "FIELDS_OUT": "fields_out", "GATHER_PUSH": "gather_push",
             "ADD_BG": "add_bg", "FIELDS_OUT_A": "fields_out_a", "FIELDS_OUT_B": "fields_out_b",
             "PARTICLES_FUSED": "particles_fused", "STATIC_FIELDS": "static_fields", "WINDOW": "window",
             "GATHER_PUSH_COORDS": "gather_push_coords", "DEPOSIT_FUSED": "deposit_fused", "COL_FWD": "col_fwd",
             "FB_IN_COL": "fb_in_col", "COL_BWD": "col_bwd", "EB_FINISH": "eb_finish"}
    assert set(ids) == set(names)
    for k, i in ids.items():
        assert PHASES[i] == names[k], (k, i, PHASES[i])
