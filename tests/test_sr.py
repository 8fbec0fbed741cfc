"""Synchrotron-radiation post-processing (reference f90/SR.f90 behind moduls/SR.py) and the utils.f90 helpers
`intens_profo` / `density_2x`.

CPU: the C++ oracle against the whole-array numpy restatement and analytic known answers.
GPU: the CUDA kernels (csrc/sr.cu) through the fimera-compatible C-ABI shim against the oracle.

Tolerance: spectra are |sum_t A e^{i omega phi}|^2 with omega*phi up to ~1e3 rad here; one ulp in phi moves a
term by eps*omega*phi ~ 2e-13, and the oscillating sum cancels by up to ~1e2, so two correct FP64
implementations agree to ~1e-11 relative (the CUDA kernel evaluates phi with the reference's operation order
and correctly rounded operations to keep it there)."""
import numpy as np
import pytest

from oracle import np_ref
from util import assert_close

SR_TOL = 1e-11


def tracks(nt, n, seed, gamma=20.0):
    """wiggling relativistic electrons: longitudinal drift + transverse oscillation, momenta consistent with it"""
    rng = np.random.default_rng(seed)
    dt = 0.05
    t = (np.arange(nt) + 1.0)[:, None] * dt
    k_u = 2 * np.pi / 3.0 * (1 + 0.05 * rng.standard_normal(n))[None, :]
    a_u = (0.8 + 0.2 * rng.random(n))[None, :]
    g = gamma * (1 + 0.02 * rng.standard_normal(n))[None, :]
    ph = 2 * np.pi * rng.random(n)[None, :]
    mom = np.zeros((3, nt, n), order="F")
    mom[2] = a_u * np.cos(k_u * t + ph)
    mom[1] = 0.1 * a_u * np.sin(k_u * t + ph) + 0.01 * rng.standard_normal(n)[None, :]
    mom[0] = np.sqrt(g ** 2 - 1 - mom[1] ** 2 - mom[2] ** 2)
    gam = np.sqrt(1 + (mom ** 2).sum(0))
    coords = np.asfortranarray(np.cumsum(mom / gam * dt, axis=1) + 0.01 * rng.standard_normal((3, 1, n)))
    mom_prv = np.asfortranarray(np.concatenate([mom[:, :1], mom[:, :-1]], axis=1))
    w = -(0.5 + rng.random(n))  # charge-signed weights: SR uses |w| (SR.f90:56)
    return coords, mom_prv, mom, w, dt


def far_grid(nom, nth, nph, gamma=20.0):
    omega = np.linspace(5.0, 2.5 * gamma ** 2 / 3.0, nom)
    theta = np.linspace(0.0, 2.0 / gamma, nth)
    phi = 2 * np.pi / nph * np.arange(nph)
    return [omega, np.sin(theta), np.cos(theta), np.sin(phi), np.cos(phi)]


FAR_CASES = [(40, 3, 7, 2, 3), (130, 5, 33, 3, 2), (300, 2, 150, 1, 1), (1, 1, 1, 1, 1)]


@pytest.mark.parametrize("nt,n,nom,nth,nph", FAR_CASES)
@pytest.mark.parametrize("comp", [0, 1, 2, 3])
def test_far_oracle_vs_numpy(ofim, nt, n, nom, nth, nph, comp):
    x, mp, mn, w, dt = tracks(nt, n, nt + n)
    g = far_grid(nom, nth, nph)
    s0 = np.asfortranarray(np.random.default_rng(1).random((nom, nth, nph)))
    if comp == 0:
        got = ofim.sr_calc_far_tot(s0.copy(order="F"), x, mp, mn, w, dt, *g)
    else:
        got = ofim.sr_calc_far_comp(s0.copy(order="F"), x, mp, mn, w, comp, dt, *g)
    assert_close(got, np_ref.sr_calc_far(s0, x, mp, mn, w, dt, *g, comp=comp), SR_TOL, "sr_calc_far")


def near_args(circ, nom, n1, n2):
    omega = np.linspace(5.0, 300.0, nom)
    if circ:
        phi = 2 * np.pi / n2 * np.arange(n2)
        return [omega, np.linspace(0.0, 3.0, n1), np.sin(phi), np.cos(phi), 40.0]
    return [omega, np.linspace(-3.0, 3.0, n1), np.linspace(-2.0, 2.0, n2), 40.0]


@pytest.mark.parametrize("circ", [False, True])
@pytest.mark.parametrize("comp", [0, 1, 2, 3])
@pytest.mark.parametrize("nt,n,nom,n1,n2", [(50, 3, 9, 3, 2), (140, 4, 40, 2, 3)])
def test_near_oracle_vs_numpy(ofim, circ, comp, nt, n, nom, n1, n2):
    x, _, m, w, dt = tracks(nt, n, 3 * nt + n)
    g = near_args(circ, nom, n1, n2)
    s0 = np.zeros((nom, n1, n2), order="F")
    name = "sr_calc_near" + ("circ" if circ else "") + ("_tot" if comp == 0 else "_comp")
    args = [s0.copy(order="F"), x, m, w] + ([comp] if comp else []) + [dt] + g
    got = getattr(ofim, name)(*args)
    if circ:
        ref = np_ref.sr_calc_near(s0, x, m, w, dt, g[0], g[1], None, g[4], comp=comp, circ=(g[2], g[3]))
    else:
        ref = np_ref.sr_calc_near(s0, x, m, w, dt, g[0], g[1], g[2], g[3], comp=comp)
    assert_close(got, ref, SR_TOL, name)


def test_far_known_answers(ofim):
    """(i) uniform motion radiates nothing; (ii) the three components add up to the total; (iii) weights enter
    as |w| and linearly; (iv) the spectrum accumulates into `spect`."""
    x, mp, mn, w, dt = tracks(60, 4, 9)
    g = far_grid(11, 2, 2)
    z = np.zeros((11, 2, 2), order="F")
    drift = np.asfortranarray(np.broadcast_to(mn[:, :1], mn.shape).copy())
    xd = np.asfortranarray(np.cumsum(drift / np.sqrt(1 + (drift ** 2).sum(0)) * dt, axis=1))
    assert np.abs(ofim.sr_calc_far_tot(z.copy(order="F"), xd, drift, drift, w, dt, *g)).max() == 0.0
    tot = ofim.sr_calc_far_tot(z.copy(order="F"), x, mp, mn, w, dt, *g)
    parts = sum(ofim.sr_calc_far_comp(z.copy(order="F"), x, mp, mn, w, c, dt, *g) for c in (1, 2, 3))
    assert_close(parts, tot, 1e-13, "sum of components")
    assert_close(ofim.sr_calc_far_tot(z.copy(order="F"), x, mp, mn, -3.0 * w, dt, *g), 3.0 * tot, 1e-13, "|w| scaling")
    twice = ofim.sr_calc_far_tot(tot.copy(order="F"), x, mp, mn, w, dt, *g)
    assert_close(twice, 2.0 * tot, 1e-13, "accumulation")
    assert np.abs(ofim.sr_calc_far_comp(z.copy(order="F"), x, mp, mn, w, 7, dt, *g)).max() == 0.0  # SR.f90:225


def test_far_first_step_guard(ofim):
    """Q13: C3_prev starts at 0 (SR.f90:65), so the first step enters only for omega*|C3(1)| < pi."""
    x, mp, mn, w, dt = tracks(1, 1, 2)
    mp = np.asfortranarray(0.9 * mp)  # a non-zero acceleration in the single step
    g = far_grid(1, 1, 1)
    c3 = 2 * np.pi * (dt - x[0, 0, 0])  # theta = 0: n = x-axis
    for om, expect_zero in ((0.5 * np.pi / abs(c3), False), (2.0 * np.pi / abs(c3), True)):
        g[0] = np.array([om])
        s = ofim.sr_calc_far_tot(np.zeros((1, 1, 1), order="F"), x, mp, mn, w, dt, *g)
        assert (s[0, 0, 0] == 0.0) == expect_zero


def test_intens_profo_and_density(ofim):
    rng = np.random.default_rng(5)
    for nm in (1, 3, 5):
        fld = np.asfortranarray(rng.standard_normal((17, 9, nm, 3)) + 1j * rng.standard_normal((17, 9, nm, 3)))
        got = ofim.intens_profo(fld, 12)
        assert got.shape == (12, 8)
        assert_close(got, np_ref.intens_profo(fld, 12), 1e-12, "intens_profo")
    # a pure m=0 field has no angular dependence
    fld = np.zeros((5, 4, 3, 3), dtype=complex, order="F")
    fld[:, :, 1, :] = rng.standard_normal((5, 4, 3))
    p = ofim.intens_profo(fld, 7)
    assert_close(p, np.broadcast_to(p[:1], p.shape), 1e-13, "m=0 isotropy")
    with pytest.raises(ofim.error):
        ofim.intens_profo(np.zeros((5, 4, 2, 3), dtype=complex, order="F"), 7)

    n = 5000
    x, y, w = rng.uniform(-1.2, 1.2, n), rng.uniform(-0.7, 2.4, n), rng.random(n)
    grid = np.array([-1.0, 1.0, -0.5, 2.0])
    d = ofim.density_2x(x, y, w, grid, 20, 13)
    assert d.shape == (25, 18)
    assert_close(d, np_ref.density_2x(x, y, w, grid, 20, 13), 1e-12, "density_2x")
    # the 5-node shape sums to one: total = sum of the weights inside the box / cell area
    inside = (x >= -1) & (x <= 1) & (y >= -0.5) & (y <= 2.0)
    assert abs(d.sum() * (2.0 / 20) * (2.5 / 13) - w[inside].sum()) < 1e-9 * w[inside].sum()


def test_sr_driver_class_call_convention(ofim):
    """the argument lists moduls/SR.py builds (Args['DepFact'], :71-73,97-98,124-126) go straight through"""
    x, mp, mn, w, dt = tracks(30, 2, 4)
    dep_far = [dt] + far_grid(6, 3, 4)
    rad = np.zeros((6, 3, 4), order="F")
    out = ofim.sr_calc_far_tot(rad, x, mp, mn, w, *dep_far)
    assert out is rad and rad.max() > 0  # intent(in,out): same object, modified in place
    out = ofim.sr_calc_far_comp(rad, x, mp, mn, w, 3, *dep_far)
    assert out is rad
    with pytest.raises(ofim.error):
        ofim.sr_calc_far_tot(rad, x, mp[:, :-1], mn, w, *dep_far)  # f2py shape check


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("nt,n,nom,nth,nph", FAR_CASES + [(700, 40, 257, 4, 3)])
@pytest.mark.parametrize("comp", [0, 1, 2, 3])
def test_gpu_far(ofim, gfim, nt, n, nom, nth, nph, comp):
    x, mp, mn, w, dt = tracks(nt, n, nt + n)
    g = far_grid(nom, nth, nph)
    s0 = np.asfortranarray(np.random.default_rng(1).random((nom, nth, nph)) * 1e-3)
    if comp == 0:
        a = ofim.sr_calc_far_tot(s0.copy(order="F"), x, mp, mn, w, dt, *g)
        b = gfim.sr_calc_far_tot(s0.copy(order="F"), x, mp, mn, w, dt, *g)
    else:
        a = ofim.sr_calc_far_comp(s0.copy(order="F"), x, mp, mn, w, comp, dt, *g)
        b = gfim.sr_calc_far_comp(s0.copy(order="F"), x, mp, mn, w, comp, dt, *g)
    assert_close(b, a, SR_TOL, "sr_calc_far comp=%d" % comp)


@pytest.mark.gpu
@pytest.mark.parametrize("circ", [False, True])
@pytest.mark.parametrize("comp", [0, 1, 2, 3])
@pytest.mark.parametrize("nt,n,nom,n1,n2", [(50, 3, 9, 3, 2), (140, 4, 40, 2, 3), (520, 30, 200, 5, 4)])
def test_gpu_near(ofim, gfim, circ, comp, nt, n, nom, n1, n2):
    x, _, m, w, dt = tracks(nt, n, 3 * nt + n)
    g = near_args(circ, nom, n1, n2)
    name = "sr_calc_near" + ("circ" if circ else "") + ("_tot" if comp == 0 else "_comp")
    s0 = np.zeros((nom, n1, n2), order="F")
    args = [x, m, w] + ([comp] if comp else []) + [dt] + g
    a = getattr(ofim, name)(s0.copy(order="F"), *args)
    b = getattr(gfim, name)(s0.copy(order="F"), *args)
    assert_close(b, a, SR_TOL, name)


@pytest.mark.gpu
def test_gpu_sr_edge_cases(ofim, gfim):
    x, mp, mn, w, dt = tracks(60, 4, 9)
    g = far_grid(11, 2, 2)
    z = np.zeros((11, 2, 2), order="F")
    assert np.abs(gfim.sr_calc_far_comp(z.copy(order="F"), x, mp, mn, w, 7, dt, *g)).max() == 0.0
    with pytest.raises(gfim.error):
        gfim.sr_calc_near_comp(np.zeros((4, 2, 2), order="F"), x, mn, w, 5, dt, *near_args(False, 4, 2, 2))
    # empty track set: spect returned untouched
    e = np.zeros((3, 60, 0), order="F")
    s = gfim.sr_calc_far_tot(np.ones((11, 2, 2), order="F"), e, e, e, np.zeros(0), dt, *g)
    assert (s == 1.0).all()


@pytest.mark.gpu
def test_gpu_intens_profo_and_density(ofim, gfim):
    rng = np.random.default_rng(5)
    for nm, no in ((1, 12), (3, 12), (5, 37)):
        fld = np.asfortranarray(rng.standard_normal((130, 19, nm, 3)) + 1j * rng.standard_normal((130, 19, nm, 3)))
        assert_close(gfim.intens_profo(fld, no), ofim.intens_profo(fld, no), 1e-12, "intens_profo")
    n = 200000
    x, y, w = rng.uniform(-1.2, 1.2, n), rng.uniform(-0.7, 2.4, n), rng.random(n)
    grid = np.array([-1.0, 1.0, -0.5, 2.0])
    assert_close(gfim.density_2x(x, y, w, grid, 20, 13), ofim.density_2x(x, y, w, grid, 20, 13), 1e-12, "density_2x")
    assert gfim.density_2x(x[:0], y[:0], w[:0], grid, 4, 4).shape == (9, 9)


# ---------------------------------------------------------------- fixture from the reference's SR class
def _golden():
    import os

    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sr.npz"))


def _replay(fim, z, key, comp):
    """the calls moduls/SR.py:165-215 made when tools/gen_golden_sr.py recorded the fixture"""
    dep = [z["%s_depfact%d" % (key, i)] for i in range(6 if key != "near" else 5)]
    dep = [float(d) if d.ndim == 0 else d for d in dep]
    rad = np.zeros_like(z[key + "_rad_all"], order="F")
    x, mp, mn, w = z["coords"], z["momenta_prv"], z["momenta_nxt"], z["weights"]
    name = {"far": "sr_calc_far", "near": "sr_calc_near", "nearcirc": "sr_calc_nearcirc"}[key] + ("_comp" if comp else "_tot")
    tr = [x, mp, mn, w] if key == "far" else [x, mn, w]
    return getattr(fim, name)(rad, *tr, *([comp] if comp else []), *dep)


@pytest.mark.parametrize("key", ["far", "near", "nearcirc"])
def test_golden_sr_class_on_oracle(ofim, key):
    z = _golden()
    assert_close(_replay(ofim, z, key, 0), z[key + "_rad_all"], 1e-13, key + " all")
    assert_close(_replay(ofim, z, key, 2), z[key + "_rad_y"], 1e-13, key + " y")


@pytest.mark.gpu
@pytest.mark.parametrize("key", ["far", "near", "nearcirc"])
def test_gpu_golden_sr_class(gfim, key):
    z = _golden()
    rad = _replay(gfim, z, key, 0)
    assert_close(rad, z[key + "_rad_all"], SR_TOL, key + " all")
    assert_close(_replay(gfim, z, key, 2), z[key + "_rad_y"], SR_TOL, key + " y")
    # the integrated diagnostic the reference derives from Rad (SR.get_energy: trapezoid in omega, sum over the screen)
    want = float(z[key + "_energy"])
    scale = want / _energy_like(z, key, z[key + "_rad_all"])
    assert abs(_energy_like(z, key, rad) * scale - want) <= 1e-6 * abs(want)


def _energy_like(z, key, rad):
    """the shape of SR.get_energy (SR.py:238-270) up to its constant factors: trapezoid over omega, weighted
    screen sum; used as a ratio against the recorded value"""
    om = z[key + "_depfact1"]
    dw = np.abs(om[1:] - om[:-1])
    if key == "far":
        wgt = z["far_depfact2"][None, :, None]  # sin(theta)
    elif key == "nearcirc":
        wgt = z["nearcirc_depfact2"][None, :, None]  # R
    else:
        wgt = 1.0
    return float((((rad[1:] + rad[:-1]) * wgt).sum(-1).sum(-1) * dw).sum())


# ---------------------------------------------------------------- physics known answer: undulator resonance
def _undulator_resonance(f):
    """One electron through `undul_analytic` with the Boris pusher and the leap-frog position update, its track fed
    to the far-field SR integral: the transverse momentum amplitude is K and the on-axis spectrum peaks at the
    undulator resonance  omega_1 = 2 gamma^2 / (lambda_u (1 + K^2/2))  -- a known answer that involves the device
    field, the push prefactor 2 pi q/m dt (species.py:64,298), the track convention of moduls/SR.py:141-147 and the
    phase convention of SR.f90:87-100 at once."""
    K, lam_u, gam, nper, dt = 0.5, 1.0, 20.0, 12, 0.02
    nt = int((nper + 2) * lam_u / dt)
    x = np.asfortranarray(np.array([[-1.0 + 1e-9], [0.0], [0.0]]))
    p = np.asfortranarray(np.array([[np.sqrt(gam ** 2 - 1)], [0.0], [0.0]]))
    xh = np.zeros_like(x)
    params = np.array([K, lam_u, 0.0, float(nper)])  # a0, lambda, X0, Lx (devices.f90:174-177)
    coords = np.zeros((3, nt, 1), order="F")
    mprv, mnxt = np.zeros_like(coords), np.zeros_like(coords)
    for it in range(nt):
        x, xh = f.push_coords(x, p, xh, dt)
        fld = f.undul_analytic(x, np.zeros((6, 1), order="F"), (it + 1) * dt, params)
        mprv[:, it, 0] = p[:, 0]
        p = f.push_velocs(p, fld, -2 * np.pi * dt)
        mnxt[:, it, 0] = p[:, 0]
        coords[:, it, 0] = x[:, 0]
    w_res = 2 * gam ** 2 / (lam_u * (1 + K ** 2 / 2))
    om = np.linspace(0.3 * w_res, 1.7 * w_res, 281)
    one, zero = np.array([1.0]), np.array([0.0])
    s = f.sr_calc_far_tot(np.zeros((om.size, 1, 1), order="F"), coords, mprv, mnxt, np.array([-1.0]), dt, om,
                          zero, one, zero, one)  # theta = 0, phi = 0
    return np.abs(mnxt[1:]).max(), K, om[np.argmax(s[:, 0, 0])], w_res, nper


def test_undulator_resonance_known_answer(ofim):
    pt, K, peak, w_res, nper = _undulator_resonance(ofim)
    assert abs(pt - K) < 2e-3 * K
    assert abs(peak / w_res - 1.0) < 0.25 / nper  # line width ~ 1/N_periods


@pytest.mark.gpu
def test_gpu_undulator_resonance_known_answer(gfim):
    pt, K, peak, w_res, nper = _undulator_resonance(gfim)
    assert abs(pt - K) < 2e-3 * K
    assert abs(peak / w_res - 1.0) < 0.25 / nper
