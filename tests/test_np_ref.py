"""CPU: the C++ oracle (oracle/chimera_oracle.cpp, loop-for-loop) against the independent vectorised
numpy restatement (oracle/np_ref.py) of the same Fortran.  Neither is the reference -- the Fortran
cannot be built here -- but two restatements in different forms agreeing to round-off is the strongest
pin available for the kernels themselves ("parity unpinned")."""
import numpy as np
import pytest

from oracle import np_ref
from util import assert_close, crandn, particles, setup

REAL = ["real_m1", "real_m2", "real_m3"]
TOL = 2e-13


@pytest.mark.parametrize("n", [1, 257, 5000])
def test_push(ofim, n):
    rng = np.random.default_rng(n)
    p = np.asfortranarray(rng.standard_normal((3, n)) * 2)
    f = np.asfortranarray(rng.standard_normal((6, n)))
    x = np.asfortranarray(rng.standard_normal((3, n)))
    assert_close(ofim.push_velocs(p.copy(order="F"), f, -0.37), np_ref.push_velocs(p, f, -0.37), TOL, "push_velocs")
    xn, xc = ofim.push_coords(x.copy(order="F"), p, np.zeros_like(x), 0.05)
    rn, rc = np_ref.push_coords(x, p, 0.05)
    assert_close(xn, rn, TOL, "coord")
    assert_close(xc, rc, TOL, "coord_cntr")


@pytest.mark.parametrize("name", REAL)
def test_deposit_and_gather(ofim, name):
    S = setup(name)
    a = S.Args
    x, p, w = particles(S, 4000, 5, inside_only=True)
    dp = a["DepProj"]
    rho = ofim.dep_dens(x, w, S.zeros_sp(), a["leftX"], *dp)
    assert_close(rho, np_ref.dep_dens(x, w, S.zeros_sp(), a["leftX"], *dp), TOL, "dep_dens")
    cur = ofim.dep_curr(x, p, w, S.zeros_sp(3), a["leftX"], *dp)
    assert_close(cur, np_ref.dep_curr(x, p, w, S.zeros_sp(3), a["leftX"], *dp), TOL, "dep_curr")
    fld = crandn(np.random.default_rng(3), S.shape_sp + (6,))
    got = ofim.proj_fld(x, w, fld, np.zeros((6, x.shape[1]), order="F"), a["leftX"], *dp)
    assert_close(got, np_ref.proj_fld(x, w, fld, np.zeros((6, x.shape[1])), a["leftX"], *dp), TOL, "proj_fld")
    assert_close(ofim.eb_correction(fld.copy(order="F")), np_ref.eb_correction(fld), TOL, "eb_correction")


@pytest.mark.parametrize("name", REAL)
def test_transforms(ofim, name):
    S = setup(name)
    a = S.Args
    rng = np.random.default_rng(11)
    vec = crandn(rng, S.shape_sp + (3,))
    kx, In = a["FBCurrIn"]
    assert_close(ofim.fb_vec_in(S.zeros_fb(3), vec, a["leftX"], kx, In), np_ref.fb_in(vec, a["leftX"], kx, In), TOL, "fb_vec_in")
    assert_close(ofim.fb_scl_in(S.zeros_fb(), vec[..., 0], a["leftX"], kx, In), np_ref.fb_in(vec[..., 0], a["leftX"], kx, In),
                 TOL, "fb_scl_in")
    e, b = crandn(rng, S.shape_fb + (6,)), crandn(rng, S.shape_fb + (3,))
    kxo, Out = a["FBout"]
    assert_close(ofim.fb_eb_out(S.zeros_sp(6), e, b, a["leftX"], kxo, Out), np_ref.fb_eb_out(e, b, a["leftX"], kxo, Out), TOL,
                 "fb_eb_out")
    assert_close(ofim.fb_vec_out(b, a["leftX"], kxo, Out), np_ref.fb_out(b, a["leftX"], kxo, Out), TOL, "fb_vec_out")


@pytest.mark.parametrize("name", ["real_m2", "real_m3"])
def test_spectral_calculus(ofim, name):
    S = setup(name)
    a = S.Args
    rng = np.random.default_rng(13)
    Dp, Dm, kx = a["FBDiff"]
    v, s = crandn(rng, S.shape_fb + (3,)), crandn(rng, S.shape_fb)
    assert_close(ofim.fb_grad(S.zeros_fb(3), s, Dp, Dm, kx), np_ref.fb_grad(s, Dp, Dm, kx), TOL, "fb_grad")
    assert_close(ofim.fb_div(S.zeros_fb(), v, Dp, Dm, kx), np_ref.fb_div(v, Dp, Dm, kx), TOL, "fb_div")
    assert_close(ofim.fb_rot(S.zeros_fb(3), v, Dp, Dm, kx), np_ref.fb_rot(v, Dp, Dm, kx), TOL, "fb_rot")
    assert_close(ofim.fb_graddiv(v.copy(order="F"), Dp, Dm, kx), np_ref.fb_graddiv(v, Dp, Dm, kx), TOL, "fb_graddiv")


def test_maxwell_family(ofim):
    S = setup("real_m2")  # SpaceCharge: 5 real coefficients
    rng = np.random.default_rng(17)
    eg, j = crandn(rng, S.shape_fb + (6,)), crandn(rng, S.shape_fb + (3,))
    g0, g1 = crandn(rng, S.shape_fb + (3,)), crandn(rng, S.shape_fb + (3,))
    got = ofim.maxwell_push_with_spchrg(eg.copy(order="F"), j, g0, g1, S.PSATD_E, S.PSATD_G)
    assert_close(got, np_ref.maxwell_push_with_spchrg(eg, j, g0, g1, S.PSATD_E, S.PSATD_G), TOL, "with_spchrg")
    a = S.Args
    got = ofim.poiss_corr(j.copy(order="F"), eg[..., :3], g0, g1, a["dt_inv"], a["PoissFact"])
    assert_close(got, np_ref.poiss_corr(j, eg[..., :3], g0, g1, a["dt_inv"], a["PoissFact"]), TOL, "poiss_corr")
    S3 = setup("real_m3")  # no space charge: 3 coefficients
    eg, j = crandn(rng, S3.shape_fb + (6,)), crandn(rng, S3.shape_fb + (3,))
    got = ofim.maxwell_push_wo_spchrg(eg.copy(order="F"), j, S3.PSATD_E, S3.PSATD_G)
    assert_close(got, np_ref.maxwell_push_wo_spchrg(eg, j, S3.PSATD_E, S3.PSATD_G), TOL, "wo_spchrg")


def test_devices(ofim):
    """every routine of devices.f90: C++ loop restatement vs whole-array numpy restatement"""
    from util import device_cases

    x, f, cases = device_cases(np.random.default_rng(21))
    for name, args in cases:
        got = getattr(ofim, name)(x, f.copy(order="F"), 0.37, *args)
        want = getattr(np_ref, name)(x, f, 0.37, *args)
        assert_close(got, want, TOL, name)
        assert np.abs(got - f).max() > 1e-3, name  # the device did something


ENV = ["env_m1", "env_m3"]


@pytest.mark.parametrize("name", ENV)
def test_envelope_deposit_and_gather(ofim, name):
    """the envelope family (grid_deps_env.f90) a second time, in whole-array numpy"""
    from util import carrier_tol

    S = setup(name)
    a = S.Args
    x, p, w = particles(S, 4000, 5, inside_only=True)
    if name == "env_m1":
        p[0] += 391.0
    dp = a["DepProj"]
    tol = carrier_tol(S, TOL)
    rho = ofim.dep_dens_env(x, w, S.zeros_sp(), a["leftX"], *dp)
    assert_close(rho, np_ref.dep_dens_env(x, w, S.zeros_sp(), a["leftX"], *dp), 20 * tol, "dep_dens_env")
    cur = ofim.dep_curr_env(x, p, w, S.zeros_sp(3), a["leftX"], *dp)
    assert_close(cur, np_ref.dep_curr_env(x, p, w, S.zeros_sp(3), a["leftX"], *dp), 20 * tol, "dep_curr_env")
    assert np.abs(cur[..., :2]).max() == 0.0 and np.abs(cur[..., 2]).max() > 0.0  # Q1
    fld = crandn(np.random.default_rng(3), S.shape_sp + (6,))
    got = ofim.proj_fld_env(x, w, fld, np.zeros((6, x.shape[1]), order="F"), a["leftX"], *dp)
    assert_close(got, np_ref.proj_fld_env(x, w, fld, np.zeros((6, x.shape[1])), a["leftX"], *dp), tol, "proj_fld_env")
    assert_close(ofim.eb_correction_env(fld.copy(order="F")), np_ref.eb_correction_env(fld), TOL, "eb_correction_env")


@pytest.mark.parametrize("name", ENV)
def test_envelope_spectral_calculus(ofim, name):
    """fb_math_env.f90 in matrix notation, per mode"""
    S = setup(name)
    a = S.Args
    rng = np.random.default_rng(13)
    Dp, Dm, kx = a["FBDiff"]
    v, s = crandn(rng, S.shape_fb + (3,)), crandn(rng, S.shape_fb)
    assert_close(ofim.fb_grad_env(S.zeros_fb(3), s, Dp, Dm, kx), np_ref.fb_grad_env(s, Dp, Dm, kx), TOL, "fb_grad_env")
    assert_close(ofim.fb_div_env(S.zeros_fb(), v, Dp, Dm, kx), np_ref.fb_div_env(v, Dp, Dm, kx), TOL, "fb_div_env")
    assert_close(ofim.fb_rot_env(S.zeros_fb(3), v, Dp, Dm, kx), np_ref.fb_rot_env(v, Dp, Dm, kx), TOL, "fb_rot_env")
    assert_close(ofim.fb_graddiv_env(v.copy(order="F"), Dp, Dm, kx), np_ref.fb_graddiv_env(v, Dp, Dm, kx), TOL, "fb_graddiv_env")
    # the complex-coefficient PSATD push of the envelope solver (maxwell_solvers.f90:62-96)
    eg, j = crandn(rng, S.shape_fb + (6,)), crandn(rng, S.shape_fb + (3,))
    got = ofim.maxwell_push_wo_spchrg(eg.copy(order="F"), j, S.PSATD_E, S.PSATD_G)
    assert_close(got, np_ref.maxwell_push_wo_spchrg(eg, j, S.PSATD_E, S.PSATD_G), TOL, "wo_spchrg (complex coefficients)")


@pytest.mark.parametrize("name", ["real_m2", "real_m3", "env_m1", "env_m3"])
@pytest.mark.parametrize("nchnk,guards", [(4, 3), (4, 0), (2, 1)])
def test_chunked_deposits(ofim, name, nchnk, guards):
    """the x-chunked variants (grid_deps_chnk.f90, grid_deps_env_chnk.f90) as "plain deposit + node mask": chunk
    edge handling through loc_left / loc_right incl. what the first and last chunk lose (Q3).  Particles are sorted
    into chunks and then shaken by up to `guards` cells, so that contributions do cross chunk borders."""
    from util import carrier_tol, chunk_sorted

    S = setup(name)
    a = S.Args
    if a["Nx"] % nchnk:
        pytest.skip("grid not divisible into %d chunks" % nchnk)
    x, p, w = particles(S, 5000, 7, inside_only=True)
    if name == "env_m1":
        p[0] += 391.0
    x, p, w, ind = chunk_sorted(S, x, p, w, ofim, nchnk)
    rng = np.random.default_rng(9)
    x[0] += a["dx"] * guards * (2 * rng.random(x.shape[1]) - 1) * 0.9  # drift since the last sort, within the guards
    x[0] = np.clip(x[0], a["leftX"] + 1e-9, a["rightX"] - a["dx"] - 1e-9)
    x = np.asfortranarray(x)
    dp = a["DepProj"]
    env = "_env" if S.env else ""
    tol = 20 * carrier_tol(S, TOL) if S.env else TOL
    rho = getattr(ofim, "dep_dens%s_chnk" % env)(x, w, S.zeros_sp(), ind, guards, a["leftX"], *dp)
    want = getattr(np_ref, "dep_dens%s_chnk" % env)(x, w, S.zeros_sp(), ind, guards, a["leftX"], *dp)
    assert_close(rho, want, tol, "dep_dens%s_chnk" % env)
    cur = getattr(ofim, "dep_curr%s_chnk" % env)(x, p, w, S.zeros_sp(3), ind, guards, a["leftX"], *dp)
    want = getattr(np_ref, "dep_curr%s_chnk" % env)(x, p, w, S.zeros_sp(3), ind, guards, a["leftX"], *dp)
    assert_close(cur, want, tol, "dep_curr%s_chnk" % env)
    # and the rule really bites: the chunked result differs from the plain one at the outer edges when guards > 0
    plain = getattr(ofim, "dep_dens%s" % env)(x, w, S.zeros_sp(), a["leftX"], *dp)
    if guards > 0:
        assert np.abs(plain - rho).max() > 0


def test_particle_tools(ofim):
    """particle_tools.f90: generation, culling, chunk binning and permutation, loop form vs array form"""
    rng = np.random.default_rng(23)
    S = setup("real_m2")
    a = S.Args
    xg, rg = a["Xgrid"][:20], a["Rgrid"][:9]
    px, pr, po = np.mgrid[1:2:2j, 1:3:3j, 1:4:4j]
    packx, packr = ((px - 0.5) / 2).ravel(), ((pr - 0.5) / 3).ravel()
    packo = np.exp(2j * np.pi * (po - 1) / 4).ravel()
    rnd = np.asfortranarray(rng.random((xg.size, rg.size)))
    ppc = packx.size
    coord = np.zeros((4, (xg.size - 1) * (rg.size - 1) * ppc), order="F")
    got, n = ofim.genparts(coord, xg, rg, rnd, packx, packr, packo)
    want = np_ref.genparts(xg, rg, rnd, packx, packr, packo)
    assert n == want.shape[1]
    assert_close(got[:, :n], want, TOL, "genparts")

    x, p, w = particles(S, 3000, 5)  # includes particles outside the domain
    lims = np.asfortranarray([a["leftX"] + 0.3, a["rightX"] - 0.2, 0.0, (0.8 * a["Rgrid"].max()) ** 2])
    idx, m = ofim.sortpartsout(x, lims)
    assert np.array_equal(idx[:m], np_ref.sortpartsout(x, lims))
    for nchnk in (1, 4, 8):
        ids, ind, out = ofim.chunk_coords_boundaries(x, lims, a["Xgrid"], nchnk)
        ids2, ind2, out2 = np_ref.chunk_coords_boundaries(x, lims, a["Xgrid"], nchnk)
        assert np.array_equal(ids, ids2) and np.array_equal(ind, ind2) and out == out2
    wz = w.copy()
    wz[::7] = 0.0
    idx, m = ofim.sortoutghosts(wz)
    assert np.array_equal(idx[:m], np_ref.sortoutghosts(wz))
    keep = np.nonzero(ids >= 0)[0].astype(np.int64)[::-1].copy()
    assert np.array_equal(ofim.align_data_vec(x.copy(order="F"), keep), np_ref.align_data(x, keep))
    assert np.array_equal(ofim.align_data_scl(w.copy(), keep), np_ref.align_data(w, keep))


def test_static_kick_family_and_filter(ofim):
    """maxwell_init_push, poiss_corr_stat, field_drift, omp_* (maxwell_solvers.f90:98-320) and fb_filtr (fb_io.f90:230)"""
    S = setup("static_m2")
    a = S.Args
    rng = np.random.default_rng(29)
    eg, j, g = crandn(rng, S.shape_fb + (6,)), crandn(rng, S.shape_fb + (3,)), crandn(rng, S.shape_fb + (3,))
    c1, c2 = S.static_coeffs(50.0)
    assert_close(ofim.maxwell_init_push(eg.copy(order="F"), j, g, c1, c2), np_ref.maxwell_init_push(eg, j, g, c1, c2), TOL,
                 "maxwell_init_push")
    beta = 50.0 / np.sqrt(1 + 50.0 ** 2)
    dtc = -1j * beta * a["kx"]
    got = ofim.poiss_corr_stat(j.copy(order="F"), eg[..., :3], g, dtc, a["PoissFact"])
    assert_close(got, np_ref.poiss_corr_stat(j, eg[..., :3], g, dtc, a["PoissFact"]), TOL, "poiss_corr_stat")
    assert_close(ofim.field_drift(eg.copy(order="F"), a["kx"], beta, a["dt"]), np_ref.field_drift(eg, a["kx"], beta, a["dt"]),
                 TOL, "field_drift")
    A = rng.random(S.shape_fb)
    assert_close(ofim.omp_mult_vec(j.copy(order="F"), A), j * A[..., None], TOL, "omp_mult_vec")
    assert_close(ofim.omp_mult_scl(j[..., 0].copy(order="F"), A), j[..., 0] * A, TOL, "omp_mult_scl")
    assert_close(ofim.omp_add_vec(j.copy(order="F"), g), j + g, TOL, "omp_add_vec")
    assert_close(ofim.omp_add_scl(j[..., 1].copy(order="F"), g[..., 2]), j[..., 1] + g[..., 2], TOL, "omp_add_scl")
    prof = S.get_damp_profile(6)
    assert_close(ofim.fb_filtr(j.copy(order="F"), a["leftX"], a["kx"], prof, 0), np_ref.fb_filtr(j, a["leftX"], a["kx"], prof, 0),
                 TOL, "fb_filtr")


def _random_setup(seed):
    """a small solver dictionary with odd sizes: grids that are not the six fixed test setups"""
    from chimera_b200.solver_setup import SolverSetup

    rng = np.random.default_rng(seed)
    nchnk = int(rng.choice([1, 2, 4]))
    nx = 2 * nchnk * int(rng.integers(3, 7))
    dx, dr = float(rng.uniform(0.03, 0.2)), float(rng.uniform(0.1, 0.4))
    nr = int(rng.integers(5, 12))
    left = float(rng.uniform(-5, 2))
    env = bool(rng.integers(0, 2))
    cfg = dict(Grid=(left, left + nx * dx, nr * dr, dx, dr), TimeStep=float(rng.uniform(0.02, 0.1)),
               MaxAzimuthMode=int(rng.integers(1, 3)), Features=("SpaceCharge",) if (not env and rng.integers(0, 2)) else ())
    if nchnk > 1:
        cfg["Xchunked"] = (nchnk, int(rng.integers(0, 3)))
    if env:
        cfg["KxShift"] = float(rng.uniform(5, 40))
    return SolverSetup(cfg)


@pytest.mark.parametrize("seed", range(12))
def test_random_grids_both_restatements_agree(ofim, seed):
    """the two restatements on randomly drawn grid sizes, spacings, mode counts, chunkings and solver families --
    nothing in either may depend on the shapes of the six fixed setups"""
    from util import carrier_tol

    S = _random_setup(1000 + seed)
    a = S.Args
    env = "_env" if S.env else ""
    rng = np.random.default_rng(seed)
    x, p, w = particles(S, 1500, seed, inside_only=True)
    dp = a["DepProj"]
    tol = 20 * carrier_tol(S, TOL) if S.env else TOL
    rho = getattr(ofim, "dep_dens" + env)(x, w, S.zeros_sp(), a["leftX"], *dp)
    assert_close(rho, getattr(np_ref, "dep_dens" + env)(x, w, S.zeros_sp(), a["leftX"], *dp), tol, "dep_dens" + env)
    cur = getattr(ofim, "dep_curr" + env)(x, p, w, S.zeros_sp(3), a["leftX"], *dp)
    assert_close(cur, getattr(np_ref, "dep_curr" + env)(x, p, w, S.zeros_sp(3), a["leftX"], *dp), tol, "dep_curr" + env)
    fld = crandn(rng, S.shape_sp + (6,))
    got = getattr(ofim, "proj_fld" + env)(x, w, fld, np.zeros((6, x.shape[1]), order="F"), a["leftX"], *dp)
    want = getattr(np_ref, "proj_fld" + env)(x, w, fld, np.zeros((6, x.shape[1])), a["leftX"], *dp)
    assert_close(got, want, carrier_tol(S, TOL), "proj_fld" + env)
    if "Xchunked" in a:
        from util import chunk_sorted

        nchnk, guards = a["Xchunked"]
        xs, ps, ws, ind = chunk_sorted(S, x, p, w, ofim, nchnk)
        got = getattr(ofim, "dep_dens%s_chnk" % env)(xs, ws, S.zeros_sp(), ind, guards, a["leftX"], *dp)
        want = getattr(np_ref, "dep_dens%s_chnk" % env)(xs, ws, S.zeros_sp(), ind, guards, a["leftX"], *dp)
        assert_close(got, want, tol, "dep_dens%s_chnk" % env)
    Dp, Dm, kx = a["FBDiff"]
    v, s = crandn(rng, S.shape_fb + (3,)), crandn(rng, S.shape_fb)
    for name, args, out in (("fb_grad", (s,), S.zeros_fb(3)), ("fb_div", (v,), S.zeros_fb()), ("fb_rot", (v,), S.zeros_fb(3))):
        got = getattr(ofim, name + env)(out, *args, Dp, Dm, kx)
        assert_close(got, getattr(np_ref, name + env)(*args, Dp, Dm, kx), TOL, name + env)
    got = getattr(ofim, "fb_graddiv" + env)(v.copy(order="F"), Dp, Dm, kx)
    assert_close(got, getattr(np_ref, "fb_graddiv" + env)(v, Dp, Dm, kx), TOL, "fb_graddiv" + env)
    if not S.env:
        vec = crandn(rng, S.shape_sp + (3,))
        kxi, In = a["FBCurrIn"]
        assert_close(ofim.fb_vec_in(S.zeros_fb(3), vec, a["leftX"], kxi, In), np_ref.fb_in(vec, a["leftX"], kxi, In), TOL, "fb_vec_in")
        kxo, Out = a["FBout"]
        assert_close(ofim.fb_vec_out(v, a["leftX"], kxo, Out), np_ref.fb_out(v, a["leftX"], kxo, Out), TOL, "fb_vec_out")
