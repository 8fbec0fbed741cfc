"""Resident mode of the per-function drop-in (chimera_b200/resident.py): numpy's allocator on CUDA managed memory, no
staging copies, the driver's whole-array statements on the device -- driven by the reference's call sequence
(tests/pic_ref.py, a restatement of ChimeraRun.make_step incl. its numpy-side mutations) and compared with the same
sequence on the oracle."""
import copy
import ctypes

import numpy as np
import pytest

from pic_ref import RefRun, RefSpecies
from util import SETUPS, TOL, assert_close, carrier_tol, match, particles, plasma, seed_fields, setup
from chimera_b200.solver_setup import SolverSetup

pytestmark = pytest.mark.gpu


@pytest.fixture()
def resident(gfim):
    from chimera_b200 import resident as R

    R.enable(threshold=4096)  # the test setups are small: put (almost) every array in managed memory
    old = R.THRESHOLD
    R.THRESHOLD = 4096
    yield R
    R.THRESHOLD = old
    R.disable()


def traffic(reset=True):
    from chimera_b200 import _lib

    h2d, d2h = ctypes.c_longlong(), ctypes.c_longlong()
    _lib.load().chimera_host_traffic(ctypes.byref(h2d), ctypes.byref(d2h), int(reset))
    return h2d.value, d2h.value


@pytest.mark.parametrize("name,ions", [("real_m2", True), ("env_m1", False), ("static_m2", False)])
def test_resident_step_sequence_matches_the_oracle(ofim, gfim, resident, name, ions):
    S = SolverSetup(copy.deepcopy(SETUPS[name]))
    x, p, w = plasma(S, 2, 2, 41)
    if name == "env_m1":
        p[0] += 391.0
    px0 = 50.0 if name == "static_m2" else 0.0
    p[0] += px0
    eg0 = seed_fields(S, 42)
    runs = []
    for fim in (ofim, gfim):
        sp = [RefSpecies(x, p, w)]
        if ions:
            xi, pi_, wi = plasma(S, 2, 2, 48)
            sp.append(RefSpecies(xi, 0 * pi_, -wi, charge=1.0, mass=1886.0, still=True))
        r = RefRun(fim, S, sp, background=ions)  # arrays created here: managed when `fim` is the CUDA drop-in's process
        if name != "static_m2":
            r.EG_fb[:] = eg0
        r.make_halfstep(px0=(px0,) * len(sp))
        for _ in range(5):  # Xchunked (4,3): includes a re-binning step (argsort + align_data on the arrays)
            r.make_step()
        runs.append(r)
    o, g = runs
    tol = carrier_tol(S, 10 * TOL)
    assert isinstance(g.EG_fb, resident.ResidentArray) and isinstance(g.sp[0].momenta, resident.ResidentArray)
    assert g.sp[0].momenta.flags.owndata  # the driver resizes its particle arrays (species.py:234-254)
    assert_close(np.asarray(g.EG_fb), o.EG_fb, tol, "EG_fb")
    assert_close(np.asarray(g.EB), o.EB, tol, "EB")
    assert_close(np.asarray(g.J), o.J, 20 * tol if S.env else tol, "J")
    perm = match(o.sp[0].weights, np.asarray(g.sp[0].weights))
    assert_close(np.asarray(g.sp[0].momenta)[:, perm], o.sp[0].momenta, tol, "momenta")
    assert_close(np.asarray(g.sp[0].coords)[:, perm], o.sp[0].coords, tol, "coords")


def test_resident_calls_move_no_bytes(gfim, resident):
    """after the first touch every argument is device-accessible: the staged-copy counters stay at zero"""
    S = SolverSetup(copy.deepcopy(SETUPS["real_m2"]))
    x, p, w = plasma(S, 2, 2, 43)
    r = RefRun(gfim, S, [RefSpecies(x, p, w)])
    r.EG_fb[:] = seed_fields(S, 44)
    r.make_halfstep()
    r.make_step()
    traffic(reset=True)
    r.make_step()
    h2d, d2h = traffic()
    # what is still copied: scalars-by-array below the threshold (chunk table, Rgrid, kx ...)
    moved_before = sum(a.nbytes for a in (r.J, r.Rho, r.EB, r.EG_fb, r.J_fb, r.sp[0].coords, r.sp[0].momenta)) * 4
    assert h2d + d2h < 0.02 * moved_before, (h2d, d2h, moved_before)


def test_host_side_access_stays_coherent(ofim, gfim, resident):
    """anything the driver (or a user) does to a returned array on the host is seen by the next call, and the other
    way round: element writes, in-place ufuncs, fancy indexing, .copy(), np.asarray, slices passed as in/out"""
    S = setup("real_m2")
    a = S.Args
    rng = np.random.default_rng(5)
    from util import crandn

    V = crandn(rng, S.shape_fb + (3,))
    go = ofim.fb_graddiv(V.copy(order="F"), *a["FBDiff"])
    gg = gfim.fb_graddiv(V.copy(order="F"), *a["FBDiff"])
    assert isinstance(gg, resident.ResidentArray)
    assert_close(gg.copy(), go, TOL, "first call")
    for arr in (go, gg):  # host-side mutations, none of them a fast path
        arr[3, 2, 1, 0] = 7.0 - 2.0j
        arr *= 0.5
        arr[::2] += 1.0
        arr[np.abs(arr) > 40.0] = 0.0
    assert_close(np.asarray(gg), go, TOL, "host-side writes on the resident array")
    go = ofim.fb_graddiv(go, *a["FBDiff"])
    gg = gfim.fb_graddiv(gg, *a["FBDiff"])
    assert_close(np.asarray(gg), go, 10 * TOL, "after host-side writes")
    # the fast paths: fill, copy, +=
    z = gfim.omp_mult_vec(gg, a["DepFact"])
    z2 = z.copy(order="F").view(resident.ResidentArray)
    z[:] = 0.0
    assert not np.asarray(z).any()
    z[:] = z2
    z += z2
    assert np.array_equal(np.asarray(z), 2 * np.asarray(z2))
    z[:] = 1.5 - 0.5j
    assert np.all(np.asarray(z) == 1.5 - 0.5j)
    # a slice of a resident array as in/out argument aliases its parent (solvers.py:548, 621-632)
    EG = crandn(rng, S.shape_fb + (6,))
    EGo = EG.copy(order="F")
    EGg = gfim.field_drift(EG.copy(order="F"), a["kx"], 0.5, a["TimeStep"])  # now resident
    EGo = ofim.field_drift(EGo, a["kx"], 0.5, a["TimeStep"])
    EGg[:, :, :, 3:] = gfim.omp_mult_vec(EGg[:, :, :, 3:], a["PoissFact"])
    EGo[:, :, :, 3:] = ofim.omp_mult_vec(EGo[:, :, :, 3:], a["PoissFact"])
    assert_close(np.asarray(EGg), EGo, TOL, "in/out slice")


def test_resident_particle_arrays_can_be_resized(ofim, gfim, resident):
    """species.py:234-254: the driver grows its particle arrays with ndarray.resize(refcheck=False) and zeroes EB"""
    S = setup("real_m2")
    a = S.Args
    x, p, w = particles(S, 3000, 7, inside_only=True)
    rng = np.random.default_rng(8)
    f = np.asfortranarray(rng.standard_normal((6, 3000)))
    pg = gfim.push_velocs(p.copy(order="F"), f, 0.3)
    po = ofim.push_velocs(p.copy(order="F"), f, 0.3)
    assert isinstance(pg, resident.ResidentArray) and pg.flags.owndata
    pg.resize((3, 3500), refcheck=False)
    po.resize((3, 3500), refcheck=False)
    assert_close(np.asarray(pg)[:, :3000], po[:, :3000], TOL, "data kept by resize")
    pg[:, 3000:] = 0.25
    po[:, 3000:] = 0.25
    f2 = np.asfortranarray(rng.standard_normal((6, 3500)))
    pg = gfim.push_velocs(pg, f2, 0.3)
    po = ofim.push_velocs(po, f2, 0.3)
    assert_close(np.asarray(pg), po, TOL, "push_velocs after resize")
    from util import crandn

    F = crandn(rng, S.shape_sp + (6,))
    eb_g = gfim.proj_fld(x, w, F, np.zeros((6, 3000), order="F"), a["leftX"], *a["DepProj"])
    eb_o = ofim.proj_fld(x, w, F, np.zeros((6, 3000), order="F"), a["leftX"], *a["DepProj"])
    eb_g[:] = 0.0  # Specie.make_field
    eb_o[:] = 0.0
    eb_g = gfim.proj_fld(x, w, F, eb_g, a["leftX"], *a["DepProj"])
    eb_o = ofim.proj_fld(x, w, F, eb_o, a["leftX"], *a["DepProj"])
    assert_close(np.asarray(eb_g), eb_o, TOL, "proj_fld into a re-zeroed resident EB")


def test_input_only_arrays_are_upgraded_where_they_are_held(ofim, gfim, resident):
    """gradRho_fb_prv is never returned by a call: the driver keeps its np.zeros object and `prv[:] = nxt`
    (chimera_main.py:110) would be a CPU copy.  The first call that sees it as intent(in) rebinds the dictionary entry
    to a ResidentArray view of the same memory."""
    from util import crandn

    S = setup("real_m2")
    a = S.Args
    rng = np.random.default_rng(9)
    data = {"prv": np.zeros(S.shape_fb + (3,), dtype=complex, order="F"), "nxt": None,
            "J": crandn(rng, S.shape_fb + (3,)), "v": crandn(rng, S.shape_fb + (3,))}
    raw = data["prv"]
    data["nxt"] = gfim.omp_mult_vec(crandn(rng, S.shape_fb + (3,)), a["DepFact"])  # resident (returned by a call)
    want = ofim.poiss_corr(data["J"].copy(order="F"), data["v"], raw.copy(order="F"), np.asarray(data["nxt"]), a["dt_inv"], a["PoissFact"])
    got = gfim.poiss_corr(data["J"].copy(order="F"), data["v"], data["prv"], data["nxt"], a["dt_inv"], a["PoissFact"])
    assert_close(np.asarray(got), want, TOL, "poiss_corr")
    assert isinstance(data["prv"], resident.ResidentArray) and data["prv"].base is raw  # same memory, upgraded in the dict
    data["prv"][:] = data["nxt"]  # now a device copy
    assert np.array_equal(raw, np.asarray(data["nxt"]))


def test_numpy_expressions_on_resident_arrays_run_on_the_device(ofim, gfim, resident):
    """chimera_main.py:121-122 (PXmean of the 'StaticKick' schedule) and the reductions of moduls/diagnostics.py on
    arrays a call returned: common ufuncs and add / max / min reductions go through torch on the same managed memory
    (strided views included); the results are what numpy gives on host copies"""
    S = setup("real_m2")
    a = S.Args
    x, p, w = particles(S, 40000, 17, inside_only=True)
    rng = np.random.default_rng(18)
    f = np.asfortranarray(rng.standard_normal((6, 40000)))
    pg = gfim.push_velocs(p.copy(order="F"), f, 0.3)
    wg = np.array(w).view(resident.ResidentArray)  # the driver's weights: managed, made resident by any call that returns it
    ph, wh = np.array(np.asarray(pg)), np.array(w)  # host copies
    got = (pg[0] * wg).sum() / wg.sum()
    want = (ph[0] * wh).sum() / wh.sum()
    assert abs(got / want - 1) < 1e-13
    assert isinstance(pg[0] * wg, resident.ResidentArray)
    from util import crandn

    EG = gfim.field_drift(crandn(rng, S.shape_fb + (6,)), a["kx"], 0.5, a["TimeStep"])
    EGh = np.array(np.asarray(EG))
    ef = np.array(a["EnergyFact"]).view(resident.ResidentArray)
    got = ((np.abs(EG[:, :, :, :3]) ** 2).sum(-1) * ef).sum(-1).sum(-1)      # Diagnostics.nrg_out, diagnostics.py:109-124
    want = ((np.abs(EGh[:, :, :, :3]) ** 2).sum(-1) * a["EnergyFact"]).sum(-1).sum(-1)
    assert_close(np.asarray(got), want, 1e-13, "nrg_out expression")
    assert abs(np.abs(EG).max() - np.abs(EGh).max()) == 0.0
    z = EG * 2.0 - EG
    assert_close(np.asarray(z), EGh, 1e-15, "element-wise chain")
    EG *= 0.5   # in place, out= path
    assert_close(np.asarray(EG), 0.5 * EGh, 1e-15, "in-place multiply")
    # unsupported pieces fall back to numpy on the same memory
    assert np.allclose(np.sin(np.asarray(pg[1, :10])), np.sin(ph[1, :10]))
    assert np.array_equal(np.asarray(np.floor(pg)), np.floor(ph))


def test_blocks_prefetched_once_stay_correct_under_unreported_host_writes(ofim, gfim, resident):
    """Managed blocks of 4 MB and more are prefetched at their first use after allocation, after host accesses the
    ResidentArray layer sees and at every 16th use, not at every call (csrc/api_host.cu managed_needs_prefetch).  What the
    host does to such an array behind the library's back -- here through plain ndarray views, which no hook sees -- must
    still be what the next call computes with (managed memory is coherent): 20 calls cross the 16th-use rule."""
    rng = np.random.default_rng(5)
    n = 400000  # (3, n) float64 = 9.6 MB: above the always-prefetch size
    x = np.zeros((3, n), order="F")  # np.zeros + a fill on the host: the static-kick tables' pattern
    x[:] = rng.standard_normal((3, n))
    p = np.asfortranarray(rng.standard_normal((3, n)))
    xc = np.zeros((3, n), order="F")
    assert resident.accessible(x) and resident.accessible(p)
    xo, po = np.array(x, order="F"), np.array(p, order="F")  # the oracle's copies (ordinary host arrays either way)
    dt = 0.05
    for k in range(20):
        x, xc = gfim.push_coords(x, p, xc, dt)
        xo, xco = ofim.push_coords(xo, po, np.zeros((3, n), order="F"), dt)
        assert_close(np.asarray(x), xo, TOL, "coords after call %d" % k)  # a stale page would be an O(1) error
        assert_close(np.asarray(xc), xco, TOL, "coords_halfstep after call %d" % k)
        if k % 3 == 0:  # an unreported host write into both inputs
            np.asarray(p)[:, k::7] *= -0.5
            po[:, k::7] *= -0.5
            np.asarray(x)[1, ::5] += 0.25
            xo[1, ::5] += 0.25
