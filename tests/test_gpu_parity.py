"""GPU parity: every C-ABI entry point of libchimera_b200.so against the CPU oracle on the same
seeded inputs (1e-12 relative L2 for floating point, exact for integer/index outputs)."""
import numpy as np
import pytest

from util import assert_close, chunk_sorted, crandn, particles, setup

pytestmark = pytest.mark.gpu

ALL = ["real_m2", "real_m3", "real_m1", "env_m1", "env_m3"]
REAL = ["real_m2", "real_m3", "real_m1"]
ENV = ["env_m1", "env_m3"]


def both(ofim, gfim, name, *args, copy_idx=()):
    """call oracle and GPU with independent copies of the in/out arguments"""
    def cp(a):
        return [x.copy(order="F") if isinstance(x, np.ndarray) else x for x in a]
    ro = getattr(ofim, name)(*cp(args))
    rg = getattr(gfim, name)(*cp(args))
    return ro, rg


# ------------------------------------------------------------------ particle_tools
@pytest.mark.parametrize("n", [0, 1, 1000, 100003])
def test_push_velocs(ofim, gfim, n):
    rng = np.random.default_rng(1)
    p = np.asfortranarray(rng.standard_normal((3, n)) * 3)
    f = np.asfortranarray(rng.standard_normal((6, n)) * 2)
    ro, rg = both(ofim, gfim, "push_velocs", p, f, 0.37)
    assert_close(rg, ro, what="push_velocs")


@pytest.mark.parametrize("n", [0, 7, 100003])
def test_push_coords(ofim, gfim, n):
    rng = np.random.default_rng(2)
    x = np.asfortranarray(rng.standard_normal((3, n)))
    p = np.asfortranarray(rng.standard_normal((3, n)) * 3)
    xc = np.zeros((3, n), order="F")
    (xo, co), (xg, cg) = both(ofim, gfim, "push_coords", x, p, xc, 0.05)
    assert_close(xg, xo, what="coords")
    assert_close(cg, co, what="coords_halfstep")


@pytest.mark.parametrize("n", [174763, 699051, 2100011, 5000003])
def test_large_pageable_buffers_are_staged_exactly(ofim, gfim, n):
    """arrays above 4 MB go through the page-locked staging ring (csrc/staging.cu: 4 MB chunks, 6 slots, host
    thread pool): one chunk plus 8 bytes (3 n 8 B = 4 MiB + 8), four chunks plus 8 bytes, a ragged tail, and many
    more chunks than slots (slot reuse) -- bit-exact pass-through both ways (align_data_vec is a pure permutation),
    then parity of a compute call on the same sizes"""
    rng = np.random.default_rng(n)
    dat = np.asfortranarray(rng.standard_normal((3, n)))
    idx = rng.permutation(n).astype(np.int64)
    got = gfim.align_data_vec(dat.copy(order="F"), idx)
    assert np.array_equal(got[:, :n], dat[:, idx])
    x = np.asfortranarray(rng.standard_normal((3, n)))
    p = np.asfortranarray(rng.standard_normal((3, n)) * 3)
    (xo, co), (xg, cg) = both(ofim, gfim, "push_coords", x, p, np.zeros((3, n), order="F"), 0.05)
    assert_close(xg, xo, what="coords (staged)")
    assert_close(cg, co, what="coords_halfstep (staged)")
    assert np.abs(xg - xo).max() <= 1e-14 * np.abs(xo).max()  # no chunk went astray: element-wise, not just in norm


def test_genparts(ofim, gfim):
    rng = np.random.default_rng(3)
    S = setup("real_m2")
    xg, rg_ = S.Args["Xgrid"][:20], S.Args["Rgrid"][:9]
    px, pr, po = np.mgrid[1:2:2j, 1:2:2j, 1:4:4j]
    px = np.asfortranarray((px.ravel() - 0.5) / 2)
    pr = np.asfortranarray((pr.ravel() - 0.5) / 2)
    po = np.asfortranarray(np.exp(2j * np.pi * (po.ravel() - 1) / 4))
    rnd = np.asfortranarray(rng.random((xg.shape[0], rg_.shape[0])))
    c0 = np.zeros((4, xg.shape[0] * rg_.shape[0] * 16), order="F")
    (co, no), (cg, ng) = both(ofim, gfim, "genparts", c0, xg, rg_, rnd, px, pr, po)
    assert no == ng and no > 0
    assert_close(cg, co, what="genparts")


def test_sortpartsout_and_ghosts(ofim, gfim):
    S = setup("real_m2")
    x, p, w = particles(S, 5000, 4)
    a = S.Args
    dom = np.asfortranarray([a["leftX"] + 0.5, a["rightX"] - 0.3, 0.0, (0.8 * a["Rgrid"].max()) ** 2])
    (io, no), (ig, ng) = both(ofim, gfim, "sortpartsout", x, dom)
    assert no == ng and 0 < no < 5000
    assert np.array_equal(io, ig)
    w[::7] = 0.0
    (io, no), (ig, ng) = both(ofim, gfim, "sortoutghosts", w)
    assert no == ng
    assert np.array_equal(io, ig)


@pytest.mark.parametrize("nchnk", [1, 4, 8])
def test_chunk_coords_boundaries(ofim, gfim, nchnk):
    S = setup("real_m2")
    x, p, w = particles(S, 20011, 5)
    a = S.Args
    dom = np.asfortranarray([a["leftX"], a["rightX"], 0.0, a["Rgrid"].max() ** 2])
    (ido, cho, goo), (idg, chg, gog) = both(ofim, gfim, "chunk_coords_boundaries", x, dom, a["Xgrid"], nchnk)
    assert goo == gog
    assert np.array_equal(cho, chg)
    assert np.array_equal(ido, idg)


def test_align_data(ofim, gfim):
    rng = np.random.default_rng(6)
    n0, n = 5000, 4200
    dat = np.asfortranarray(rng.standard_normal((3, n0)))
    scl = np.asfortranarray(rng.standard_normal(n0))
    idx = rng.permutation(n0)[:n]
    ro, rg = both(ofim, gfim, "align_data_vec", dat, idx)
    assert np.array_equal(ro[:, :n], rg[:, :n])
    ro, rg = both(ofim, gfim, "align_data_scl", scl, idx.astype(np.int32))  # f2py casts int32 -> int64
    assert np.array_equal(ro[:n], rg[:n])


# ------------------------------------------------------------------ deposit / gather
@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("n", [0, 64, 30011])
def test_deposit_plain(ofim, gfim, name, n):
    S = setup(name)
    a = S.Args
    x, p, w = particles(S, n, 7)
    sfx = "_env" if S.env else ""
    J = crandn(np.random.default_rng(8), S.shape_sp + (3,)) * 1e-4
    ro, rg = both(ofim, gfim, "dep_curr" + sfx, x, p, w, J, a["leftX"], *a["DepProj"])
    assert_close(rg, ro, what="dep_curr" + sfx)
    R = crandn(np.random.default_rng(9), S.shape_sp) * 1e-4
    ro, rg = both(ofim, gfim, "dep_dens" + sfx, x, w, R, a["leftX"], *a["DepProj"])
    assert_close(rg, ro, what="dep_dens" + sfx)


@pytest.mark.parametrize("name", ["real_m2", "env_m1", "env_m3"])
@pytest.mark.parametrize("guards", [0, 3])
@pytest.mark.parametrize("n", [20000, 150000])  # 150000: above the size where api_host.cu takes the CTA-binned kernel
def test_deposit_chunked(ofim, gfim, name, guards, n):
    S = setup(name)
    a = S.Args
    nchnk = 4
    x, p, w = particles(S, n, 10, inside_only=True)
    x, p, w, chunks = chunk_sorted(S, x, p, w, ofim, nchnk)
    # let particles drift up to `guards` cells after the sort, as between two sorts of the driver
    x[0] += a["dx"] * guards * (np.random.default_rng(11).random(x.shape[1]) - 0.5) * 1.9
    sfx = "_env" if S.env else ""
    J = S.zeros_sp(3)
    ro, rg = both(ofim, gfim, "dep_curr" + sfx + "_chnk", x, p, w, J, chunks, guards, a["leftX"], *a["DepProj"])
    assert_close(rg, ro, what="dep_curr" + sfx + "_chnk")
    R = S.zeros_sp()
    ro, rg = both(ofim, gfim, "dep_dens" + sfx + "_chnk", x, w, R, chunks, guards, a["leftX"], *a["DepProj"])
    assert_close(rg, ro, what="dep_dens" + sfx + "_chnk")


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("n", [0, 64, 30011, 150001])  # 150001: above the size where api_host.cu takes the CTA-binned kernel
def test_proj_fld(ofim, gfim, name, n):
    S = setup(name)
    a = S.Args
    x, p, w = particles(S, n, 12)
    F = crandn(np.random.default_rng(13), S.shape_sp + (6,))
    out = np.asfortranarray(np.random.default_rng(14).standard_normal((6, n)))
    ro, rg = both(ofim, gfim, "proj_fld" + ("_env" if S.env else ""), x, w, F, out, a["leftX"], *a["DepProj"])
    assert_close(rg, ro, what="proj_fld")


@pytest.mark.parametrize("name", ALL)
def test_eb_correction(ofim, gfim, name):
    S = setup(name)
    F = crandn(np.random.default_rng(15), S.shape_sp + (6,))
    ro, rg = both(ofim, gfim, "eb_correction" + ("_env" if S.env else ""), F)
    assert_close(rg, ro, what="eb_correction")


def test_devices(ofim, gfim):
    """every routine of devices.f90 (NEXT-1): undulators (analytic, tapered, mapped), plane wave, Gaussian packet"""
    from util import device_cases

    x, f, cases = device_cases(np.random.default_rng(16), n=5003)
    for name, args in cases:
        ro, rg = both(ofim, gfim, name, x, f, 0.3, *args)
        assert_close(rg, ro, what=name)
    # empty input and f2py shape errors
    e = gfim.planewave(np.zeros((3, 0), order="F"), np.zeros((6, 0), order="F"), 0.0, cases[4][1][0])
    assert e.shape == (6, 0)
    with pytest.raises(gfim.error):
        gfim.planewave(x, f, 0.0, np.zeros(4))


# ------------------------------------------------------------------ DHT + FFT
@pytest.mark.parametrize("name", ALL)
def test_fb_in(ofim, gfim, name):
    S = setup(name)
    a = S.Args
    rng = np.random.default_rng(17)
    J = crandn(rng, S.shape_sp + (3,))
    ro, rg = both(ofim, gfim, "fb_vec_in", S.zeros_fb(3), J, a["leftX"], *a["FBCurrIn"])
    assert_close(rg, ro, what="fb_vec_in")
    R = crandn(rng, S.shape_sp)
    ro, rg = both(ofim, gfim, "fb_scl_in", S.zeros_fb(), R, a["leftX"], *a["FBIn"])
    assert_close(rg, ro, what="fb_scl_in")


@pytest.mark.parametrize("name", ALL)
def test_fb_out(ofim, gfim, name):
    S = setup(name)
    a = S.Args
    rng = np.random.default_rng(18)
    V = crandn(rng, S.shape_fb + (3,))
    ro, rg = both(ofim, gfim, "fb_vec_out", V, a["leftX"], *a["FBout"])
    assert_close(rg, ro, what="fb_vec_out")
    Sc = crandn(rng, S.shape_fb)
    ro, rg = both(ofim, gfim, "fb_scl_out", Sc, a["leftX"], *a["FBout"])
    assert_close(rg, ro, what="fb_scl_out")
    EG = crandn(rng, S.shape_fb + (6,))
    B = crandn(rng, S.shape_fb + (3,))
    ro, rg = both(ofim, gfim, "fb_eb_out", S.zeros_sp(6), EG, B, a["leftX"], *a["FBout"])
    assert_close(rg, ro, what="fb_eb_out")
    # slices of a larger array, as the driver passes them (solvers.py:548)
    ro, rg = both(ofim, gfim, "fb_vec_out", EG[:, :, :, 3:], a["leftX"], *a["FBout"])
    assert_close(rg, ro, what="fb_vec_out(slice)")


@pytest.mark.parametrize("name", ["real_m2", "env_m1"])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_fb_filtr(ofim, gfim, name, mode):
    S = setup(name)
    a = S.Args
    V = crandn(np.random.default_rng(19), S.shape_fb + (3,))
    nf = 12
    g = np.arange(nf)
    filt = (g >= 0.75 * nf) * (0.5 - 0.5 * np.cos(np.pi * (g - 0.75 * nf) / (0.25 * nf))) ** 2
    ro, rg = both(ofim, gfim, "fb_filtr", V, a["leftX"], a["kx"], filt, mode)
    assert_close(rg, ro, what="fb_filtr")


# ------------------------------------------------------------------ spectral vector calculus
@pytest.mark.parametrize("name", ALL)
def test_fb_math(ofim, gfim, name):
    S = setup(name)
    a = S.Args
    sfx = "_env" if S.env else ""
    rng = np.random.default_rng(20)
    V = crandn(rng, S.shape_fb + (3,))
    Sc = crandn(rng, S.shape_fb)
    ro, rg = both(ofim, gfim, "fb_rot" + sfx, S.zeros_fb(3), V, *a["FBDiff"])
    assert_close(rg, ro, what="fb_rot" + sfx)
    ro, rg = both(ofim, gfim, "fb_grad" + sfx, S.zeros_fb(3), Sc, *a["FBDiff"])
    assert_close(rg, ro, what="fb_grad" + sfx)
    ro, rg = both(ofim, gfim, "fb_div" + sfx, S.zeros_fb(), V, *a["FBDiff"])
    assert_close(rg, ro, what="fb_div" + sfx)
    ro, rg = both(ofim, gfim, "fb_graddiv" + sfx, V, *a["FBDiff"])
    assert_close(rg, ro, what="fb_graddiv" + sfx)


# ------------------------------------------------------------------ PSATD elementwise family
@pytest.mark.parametrize("name", ALL)
def test_maxwell_family(ofim, gfim, name):
    S = setup(name)
    a = S.Args
    rng = np.random.default_rng(21)
    EG = crandn(rng, S.shape_fb + (6,))
    J = crandn(rng, S.shape_fb + (3,))
    g1 = crandn(rng, S.shape_fb + (3,))
    g2 = crandn(rng, S.shape_fb + (3,))
    if S.space_charge:
        ro, rg = both(ofim, gfim, "maxwell_push_with_spchrg", EG, J, g1, g2, S.PSATD_E, S.PSATD_G)
        assert_close(rg, ro, what="maxwell_push_with_spchrg")
    else:
        ro, rg = both(ofim, gfim, "maxwell_push_wo_spchrg", EG, J, S.PSATD_E, S.PSATD_G)
        assert_close(rg, ro, what="maxwell_push_wo_spchrg")
    c1, c2 = S.static_coeffs(50.0)
    ro, rg = both(ofim, gfim, "maxwell_init_push", EG, J, g1, c1, c2)
    assert_close(rg, ro, what="maxwell_init_push")
    ro, rg = both(ofim, gfim, "poiss_corr", J, g1, g2, EG[..., :3], a["dt_inv"], a["PoissFact"])
    assert_close(rg, ro, what="poiss_corr")
    DT = -1j * 0.9998 * a["kx"]
    ro, rg = both(ofim, gfim, "poiss_corr_stat", J, g1, g2, DT, a["PoissFact"])
    assert_close(rg, ro, what="poiss_corr_stat")
    ro, rg = both(ofim, gfim, "field_drift", EG, a["kx"], 0.9998, a["TimeStep"])
    assert_close(rg, ro, what="field_drift")
    ro, rg = both(ofim, gfim, "omp_mult_vec", J, a["DepFact"])
    assert_close(rg, ro, what="omp_mult_vec")
    ro, rg = both(ofim, gfim, "omp_mult_scl", J[..., 0], a["DepFact"])
    assert_close(rg, ro, what="omp_mult_scl")
    ro, rg = both(ofim, gfim, "omp_add_vec", J, g1)
    assert_close(rg, ro, what="omp_add_vec")
    ro, rg = both(ofim, gfim, "omp_add_scl", J[..., 1], g1[..., 2])
    assert_close(rg, ro, what="omp_add_scl")


# ------------------------------------------------------------------ larger DHT (several GEMM tiles, ragged edges)
def test_fb_roundtrip_large(ofim, gfim):
    from chimera_b200.solver_setup import SolverSetup
    S = SolverSetup(dict(Grid=(-8.0, 1.0, 30.0, 0.03, 0.2), TimeStep=0.03, MaxAzimuthMode=1, Features=()))
    a = S.Args
    rng = np.random.default_rng(22)
    V = crandn(rng, S.shape_sp + (3,))
    V[:, 0] = 0
    ro, rg = both(ofim, gfim, "fb_vec_in", S.zeros_fb(3), V, a["leftX"], *a["FBIn"])
    assert_close(rg, ro, what="fb_vec_in large")
    back = gfim.fb_vec_out(rg, a["leftX"], *a["FBout"]) / a["Nx"]
    assert_close(back[:, 1:], V[:, 1:], tol=1e-9, what="DHT+FFT round trip")


def test_gaussian_beam_diffraction_on_gpu(gfim):
    """the physics known answer of tests/test_oracle_kat.py through the CUDA drop-in"""
    from test_oracle_kat import gaussian_beam_diffraction

    for s, t, got, want in gaussian_beam_diffraction(gfim):
        assert abs(s - t) <= 0.051
        assert abs(got / want - 1) < 0.05, (s, got, want)
