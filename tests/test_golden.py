"""Golden fixtures recorded from the reference's own, unmodified Python driver (tools/gen_golden.py:
``Solver`` / ``Specie`` / ``ChimeraRun`` imported from the reference checkout, CPU oracle underneath).

CPU (-m "not gpu")
  * the host-side table builder (chimera_b200/solver_setup.py) reproduces the reference ``Solver``'s
    DHT / mode-coupling / PSATD tables slot for slot (reference moduls/solvers.py:27-279, 717-759);
  * the compact step sequence used everywhere in the tests (tests/pic_ref.py) replays what
    ``ChimeraRun.__init__`` + ``make_step`` (chimera_main.py:61-92) did, on the oracle.
GPU (-m gpu)
  * the device-resident engine and the host-buffer drop-in reproduce the same recorded states.

Tolerance: 1e-12 relative L2 on fields and momenta after one step is the north_star bar; the fixtures
hold 4 steps, so the GPU comparisons use 1e-11 (round-off grows with the step count)."""
import copy
import json
import os

import numpy as np
import pytest

from pic_ref import RefRun, RefSpecies
from util import SETUPS, assert_close, carrier_tol, match
from chimera_b200.solver_setup import SolverSetup

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["real_m2", "real_m3", "env_m1", "env_m3", "static_m2", "env_m1_win"]
ENGINE_NAMES = NAMES


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["cfg"]))
    return z, meta["case"], meta["nsteps"]


@pytest.mark.parametrize("name", NAMES)
def test_solver_tables_match_reference_solver(name):
    z, case, _ = load(name)
    S = SolverSetup(copy.deepcopy(SETUPS[case["setup"]]))
    a = S.Args
    for k in ("In", "InCurr", "Out", "DpS2S", "DmS2S", "DepFact", "PoissFact", "kx", "kx_env", "Rgrid", "Xgrid", "VGrid"):
        want = z["tab_" + k]
        got = np.asarray(a[k])
        assert got.shape == want.shape, (k, got.shape, want.shape)
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-13 * np.abs(want).max(), err_msg=k)
    for k, got in (("PSATD_E", S.PSATD_E), ("PSATD_G", S.PSATD_G)):
        want = z["tab_" + k]
        assert got.dtype == want.dtype and got.shape == want.shape, (k, got.dtype, want.dtype)
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-13 * np.abs(want).max(), err_msg=k)
    assert float(a["DepProj"][1]) == float(z["tab_DepProj1"]) and float(a["DepProj"][2]) == float(z["tab_DepProj2"])


def _ref_run(fim, z, case):
    S = SolverSetup(copy.deepcopy(SETUPS[case["setup"]]))
    dev = None
    if case["und"]:
        u = case["und"]
        dev = (fim.undul_analytic, [u["a0"], u["lam"], u["X0"], u["Lx"]])
    sp = [RefSpecies(z["in_coords"], z["in_momenta"], z["in_weights"], device=dev)]
    if case["ions"]:
        xi = z["in_ion_coords"]
        sp.append(RefSpecies(xi, 0 * xi, z["in_ion_weights"], charge=1.0, mass=1886.0, still=True))
    run = RefRun(fim, S, sp, background=False)
    run.EG_fb[:] = z["in_EG_fb"]
    if case.get("window"):
        run.window = _window_shifts(case, S)
    return S, run


def _window_shifts(case, S):
    """ChimeraRun.init_Moving_Frames (chimera_main.py:40-51) for a frame with 'Steps': 1, 'TimeStep': dt"""
    v, dt = case["window"]["Velocity"], S.Args["dt"]
    return (0.5 * v * dt, 0.5 * v * dt) if case["window"]["Staged"] else (v * dt, 0.0)


def _check_particles(z, prefix, x, xh, p, w, tol):
    perm = match(z[prefix + "_weights"], w)
    assert_close(p[:, perm], z[prefix + "_momenta"], tol, prefix + " momenta")
    assert_close(x[:, perm], z[prefix + "_coords"], tol, prefix + " coords")
    assert_close(xh[:, perm], z[prefix + "_coords_halfstep"], tol, prefix + " coords_halfstep")


def _replay(fim, name, tol):
    z, case, nsteps = load(name)
    S, run = _ref_run(fim, z, case)
    tol = carrier_tol(S, tol)
    # the reference applies the static kick once PER SPECIES, still ones included (chimera_main.py:74-76)
    run.make_halfstep(px0=(0.0,) * len(run.sp))
    assert_close(run.EG_fb, z["h_EG_fb"], tol, "EG_fb after make_halfstep")
    s = run.sp[0]
    _check_particles(z, "h", s.coords, s.coords_halfstep, s.momenta, s.weights, tol)
    if "h_chunks" in z.files:
        assert np.array_equal(s.chunks, z["h_chunks"])
    for _ in range(nsteps):
        if case["ions"]:
            run.dep_bg()  # frame_act -> postframe_corr -> dep_bg on every step (chimera_main.py:277-304)
        run.make_step()
    env = S.env
    for k, got in (("J", run.J), ("Rho", run.Rho), ("BckGrndRho", run.Bck), ("EB", run.EB), ("EG_fb", run.EG_fb)):
        if "s_" + k not in z.files:  # Rho / BckGrndRho exist only for a SpaceCharge solver (solvers.py:196-212)
            continue
        assert_close(got, z["s_" + k], (20 * tol) if (env and k == "J") else tol, k)
    _check_particles(z, "s", s.coords, s.coords_halfstep, s.momenta, s.weights, tol)
    if "s_chunks" in z.files:
        assert np.array_equal(s.chunks, z["s_chunks"])


@pytest.mark.parametrize("name", NAMES)
def test_step_sequence_replays_reference_driver_on_oracle(ofim, name):
    """tests/pic_ref.py + oracle == reference ChimeraRun + oracle (summation order inside a chunk is
    the only freedom: numpy argsort is unstable in the reference, species.py:382)."""
    _replay(ofim, name, 5e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_dropin_replays_golden(gfim, name):
    """chimera_b200.fimera (CUDA, host buffers) under the same sequence."""
    _replay(gfim, name, 1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ENGINE_NAMES)
def test_engine_replays_golden(gfim, name):
    from chimera_b200.engine import Engine

    tol = 1e-11
    z, case, nsteps = load(name)
    S = SolverSetup(copy.deepcopy(SETUPS[case["setup"]]))
    und = None
    if case["und"]:
        u = case["und"]
        und = {"a0": u["a0"], "lambda": u["lam"], "X0": u["X0"], "Lx": u["Lx"]}
    tol = carrier_tol(S, tol)
    eng = Engine(S, undulator=und)
    eng.add_species(z["in_coords"], z["in_momenta"], z["in_weights"])
    if case["ions"]:
        xi = z["in_ion_coords"]
        eng.add_species(xi, 0 * xi, z["in_ion_weights"], charge=1.0, mass=1886.0, still=True)
    eng.upload("EG_fb", z["in_EG_fb"])
    if case.get("window"):
        eng.set_window(case["window"]["Velocity"], staged=case["window"]["Staged"])
    eng.make_halfstep(px0=(0.0,) * (2 if case["ions"] else 1), background=False)
    assert_close(eng.download("EG_fb"), z["h_EG_fb"], tol, "EG_fb after make_halfstep")
    x, xh, p, w = eng.particles(0)
    _check_particles(z, "h", x, xh, p, w, tol)
    if case["ions"]:
        eng.deposit_background()
    eng.step(nsteps)
    for k in ("J", "Rho", "BckGrndRho", "EB", "EG_fb"):
        if "s_" + k not in z.files:
            continue
        assert_close(eng.download(k), z["s_" + k], (20 * tol) if (S.env and k == "J") else tol, k)
    x, xh, p, w = eng.particles(0)
    _check_particles(z, "s", x, xh, p, w, tol)
    if "s_chunks" in z.files:
        assert np.array_equal(eng.chunks(0), z["s_chunks"])
    eng.close()


# ------------------------------------------------------------------ LPA stage with its moving window
def _check_particles_by_position(z, prefix, x, xh, p, w, tol):
    """gen_parts gives every particle of a radial row the same weight, so particles are identified by position
    (nearest neighbour in 3-D; the match has to be one to one)"""
    from scipy.spatial import cKDTree

    want = z[prefix + "_coords"]
    assert x.shape == want.shape, (x.shape, want.shape)
    dist, perm = cKDTree(x.T).query(want.T)
    assert np.unique(perm).size == perm.size and dist.max() <= 1e-9
    assert_close(x[:, perm], want, tol, prefix + " coords")
    assert_close(xh[:, perm], z[prefix + "_coords_halfstep"], tol, prefix + " coords_halfstep")
    assert_close(p[:, perm], z[prefix + "_momenta"], tol, prefix + " momenta")
    assert np.array_equal(w[perm], z[prefix + "_weights"])


def _lpa_fixture():
    z = np.load(os.path.join(GOLDEN, "real_m2_lpa.npz"))
    meta = json.loads(str(z["cfg"]))
    case = meta["case"]
    adds = {}
    for k, (st, si) in enumerate(zip(z["add_steps"], z["add_species"])):
        adds.setdefault(int(st), {})[int(si)] = (z["add%d_coords" % k], z["add%d_momenta" % k], z["add%d_weights" % k])
    wind = {"shiftX": float(z["shiftX"]), "AbsorbLayer": case["wind"]["AbsorbLayer"], "Steps": case["wind"]["Steps"]}
    return z, case, meta["nsteps"], adds, wind


def test_lpa_window_sequence_replays_reference_driver_on_oracle(ofim):
    """doc/tests/lpa-testrun.py in miniature, recorded from the reference's own ChimeraRun (tools/gen_golden.py
    generate_lpa): species that start EMPTY, a window acting every 3 steps (damp_fields, move_frame, add_plasma with
    'IonsOnTop', damp_plasma, postframe_corr -- chimera_main.py:250-304).  tests/pic_ref.RefRun.frame_act + make_step
    on the oracle reproduce the recorded states; the particles the driver's gen_parts produced are replayed."""
    z, case, nsteps, adds, wind = _lpa_fixture()
    S = SolverSetup(copy.deepcopy(SETUPS[case["setup"]]))
    e0 = np.zeros((3, 0), order="F")
    sp = [RefSpecies(e0, e0, np.zeros(0)), RefSpecies(e0, e0, np.zeros(0), charge=1.0, mass=1886.0, still=True)]
    run = RefRun(ofim, S, sp, sort_every=0, background=True)  # species feature 'NoSorting': only the window re-bins
    run.EG_fb[:] = z["in_EG_fb"]
    run.make_halfstep(px0=(0.0, 0.0))
    assert_close(run.EG_fb, z["h_EG_fb"], 5e-13, "EG_fb after make_halfstep (no particles)")
    for i in range(1, nsteps + 1):
        if i % wind["Steps"] == 0:
            run.frame_act(wind, add=adds.get(i))
        run.make_step()
        if "s%d_EG_fb" % i in z.files:
            pre = "s%d" % i
            for k, got in (("J", run.J), ("Rho", run.Rho), ("BckGrndRho", run.Bck), ("EB", run.EB), ("EG_fb", run.EG_fb)):
                assert_close(got, z[pre + "_" + k], 5e-12, "%s at step %d" % (k, i))
            s = run.sp[0]
            _check_particles_by_position(z, pre, s.coords, s.coords_halfstep, s.momenta, s.weights, 5e-12)
            assert run.sp[1].weights.size == z[pre + "_ion_weights"].size


@pytest.mark.gpu
def test_engine_replays_lpa_window_golden(gfim):
    """the same recorded LPA-window run on the device-resident engine: Engine.frame_act between step() calls"""
    from chimera_b200.engine import Engine

    z, case, nsteps, adds, wind = _lpa_fixture()
    S = SolverSetup(copy.deepcopy(SETUPS[case["setup"]]))
    e0 = np.zeros((3, 0), order="F")
    eng = Engine(S, sort_every=0)
    eng.add_species(e0, e0, np.zeros(0), capacity=4096)
    eng.add_species(e0, e0, np.zeros(0), charge=1.0, mass=1886.0, still=True, capacity=4096)
    eng.upload("EG_fb", z["in_EG_fb"])
    eng.make_halfstep(px0=(0.0, 0.0), background=True)
    assert_close(eng.download("EG_fb"), z["h_EG_fb"], 1e-11, "EG_fb after make_halfstep (no particles)")
    for i in range(1, nsteps + 1):
        if i % wind["Steps"] == 0:
            eng.frame_act(wind, add=adds.get(i), background=True)
        eng.step(1)
        if "s%d_EG_fb" % i in z.files:
            pre = "s%d" % i
            for k in ("J", "Rho", "BckGrndRho", "EB", "EG_fb"):
                assert_close(eng.download(k), z[pre + "_" + k], 2e-11, "%s at step %d" % (k, i))
            x, xh, p, w = eng.particles(0)
            _check_particles_by_position(z, pre, x, xh, p, w, 2e-11)
            assert eng.count(1) == z[pre + "_ion_weights"].size
    eng.close()


# ------------------------------------------------------------------ the reference's Diagnostics on the FEL fixture
PWR_FCTR = 0.5 * 0.511e6 * 1.6022e-19 / 2.818e-13 * 2.9979e10  # moduls/diagnostics.py:24
NTHETA = 60  # moduls/diagnostics.py:25


def _diagnostics_like_reference(fim, S, eg_fb, left_x):
    """Diagnostics.nrg_out / pwr_out(..., 'Spot') (diagnostics.py:109-149) written against a fimera backend"""
    a = S.Args
    nrg = ((np.abs(eg_fb[:, :, :, :3]) ** 2).sum(-1) * a["EnergyFact"]).sum(-1).sum(-1)
    nrg = np.r_[nrg[nrg.shape[0] // 2 + 1:], nrg[:nrg.shape[0] // 2 + 1]]
    dat = fim.fb_vec_out(np.asfortranarray(eg_fb[:, :, :, :3]), left_x, *a["FBoutFull"])
    pwr = PWR_FCTR * 2 * a["dr"] * ((np.abs(dat) ** 2).sum(-1).sum(-1) * a["RgridFull"][None, :]).sum(-1)
    return nrg, pwr, fim.intens_profo(dat, NTHETA)


def _final_left_x(S, case, nsteps):
    lx = S.Args["leftX"]
    for _ in range(nsteps):
        for sft in _window_shifts(case, S):
            lx = lx + sft
    return lx


def test_reference_diagnostics_on_oracle(ofim):
    """the integrated diagnostics of the reference's own Diagnostics class, recorded on the FEL-window fixture, from
    the replayed state: field energy per kx, power per x, azimuthal spot profile, beam envelopes"""
    z, case, nsteps = load("env_m1_win")
    S, run = _ref_run(ofim, z, case)
    run.make_halfstep(px0=(0.0,))
    for _ in range(nsteps):
        run.make_step()
    nrg, pwr, spot = _diagnostics_like_reference(ofim, S, run.EG_fb, S.Args["leftX"])
    tol = carrier_tol(S, 1e-12)
    assert_close(nrg, z["d_nrg"], tol, "nrg_out")
    assert_close(pwr, z["d_pwr"], tol, "pwr_out")
    assert_close(spot, z["d_spot"], tol, "pwr_out spot (intens_profo)")


@pytest.mark.gpu
def test_reference_diagnostics_on_engine(gfim):
    """the same recorded diagnostics from the device: Engine.nrg_out / get_beam_envelops (reductions on the GPU) and
    fb_vec_out + intens_profo through the drop-in; north_star bar for integrated diagnostics: 1e-6"""
    from chimera_b200.engine import Engine

    z, case, nsteps = load("env_m1_win")
    S = SolverSetup(copy.deepcopy(SETUPS[case["setup"]]))
    u = case["und"]
    eng = Engine(S, undulator={"a0": u["a0"], "lambda": u["lam"], "X0": u["X0"], "Lx": u["Lx"]})
    eng.add_species(z["in_coords"], z["in_momenta"], z["in_weights"])
    eng.upload("EG_fb", z["in_EG_fb"])
    eng.set_window(case["window"]["Velocity"], staged=case["window"]["Staged"])
    eng.make_halfstep(px0=(0.0,))
    eng.step(nsteps)
    assert_close(np.asarray(eng.nrg_out()), z["d_nrg"], 1e-9, "Engine.nrg_out")
    env = eng.get_beam_envelops()
    assert np.abs(env - z["d_env"]).max() <= 1e-6 * np.abs(z["d_env"]).max(), (env, z["d_env"])
    _, pwr, spot = _diagnostics_like_reference(gfim, S, eng.download("EG_fb"), _final_left_x(S, case, nsteps))
    assert_close(pwr, z["d_pwr"], 1e-9, "pwr_out through the drop-in")
    assert_close(spot, z["d_spot"], 1e-9, "spot profile through the drop-in")
    eng.close()


# ------------------------------------------------------------------ laser injection
@pytest.mark.parametrize("name", ["real_m2", "env_m1", "env_m3"])
def test_add_gauss_beam_matches_reference_solver(ofim, name):
    """SolverSetup.add_gauss_beam against the EG_fb the reference's own Solver.add_gauss_beam produced
    (tools/gen_golden_laser.py): real solver (scalar seed through fb_scl_in, sign(kx) propagator), envelope solver
    (analytic spectral seed with the Bessel normalisation), divergence cleaning, focus propagation"""
    z = np.load(os.path.join(GOLDEN, "laser.npz"))
    S = SolverSetup(copy.deepcopy(SETUPS[name]))
    laser = dict(zip(("a0", "k0", "x0", "x_foc", "Lx", "LR"), z[name + "_laser"]))
    got = S.add_gauss_beam(ofim, laser)
    assert_close(got, z[name + "_EG_fb"], 1e-12, "EG_fb after add_gauss_beam")
    twice = S.add_gauss_beam(ofim, laser, EG_fb=got.copy(order="F"))
    assert_close(twice, 2 * z[name + "_EG_fb"], 1e-12, "accumulation into a given EG_fb")
