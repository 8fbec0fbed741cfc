"""Generate the golden fixtures under tests/golden/ (BUILD CONTAINER ONLY: needs /root/reference).

What pins what
--------------
The reference's hot path is Fortran (f90/*.f90) that cannot be compiled in this image (no gfortran,
no FFTW3), and the reference ships no golden vectors (doc/tests/*.py assert the exit code only).
What CAN be run here is the reference's *unmodified Python*: ``Solver`` (moduls/solvers.py -- builds
every DHT / mode-coupling / PSATD table the kernels consume), ``Specie`` (moduls/species.py) and
``ChimeraRun`` (moduls/chimera_main.py -- the per-step call sequence).  This script imports them from
/root/reference (tools/ref_driver.py; nothing is copied) with the CPU oracle installed as
``chimera.moduls.fimera`` and records, per configuration:

  * ``tab_*``   the operator / coefficient tables of the reference's own ``Solver``
                (pins chimera_b200/solver_setup.py, slot for slot, on the GPU box too);
  * ``in_*``    seeded particles and a seeded initial field;
  * ``h_*``     the state after ``ChimeraRun.__init__`` (= make_halfstep, chimera_main.py:61-80);
  * ``s_*``     the state after NSTEPS x ``ChimeraRun.make_step`` (chimera_main.py:82-92).

So the *sequence* (which kernel is called when, with which arrays, including the driver's numpy-side
mutations) is the reference's own; the *kernels underneath* are the oracle restatement -- the Fortran
itself stays unpinned ("parity unpinned", DESIGN.md section Oracle).  tests/test_golden.py replays the
fixtures on the oracle through tests/pic_ref.py (CPU) and on the CUDA engine and the CUDA drop-in
(GPU).

  python tools/gen_golden.py [case ...]  # rewrites tests/golden/<case>.npz (all cases without arguments)
"""
import copy
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

NSTEPS = 4

# name -> (solver setup in tests/util.SETUPS, still ions?, undulator params or None, field amplitude)
CASES = {
    "real_m2": dict(setup="real_m2", ions=True, und=None, amp=0.5),     # LPA-like: space charge, background, chunked
    "real_m3": dict(setup="real_m3", ions=False, und=None, amp=0.5),    # 3 modes, no space charge, not chunked
    # FEL-like: a gamma = 391 beam.  Slow particles would make this setup a round-off amplifier (a position
    # error dx changes the gathered carrier phase by kx0 dx ~ 7e5 dx, the push feeds it back into x: x4400 per
    # step for v << c, x1e-4 for gamma = 391 where dv = dp / gamma^3)
    "env_m1": dict(setup="env_m1", ions=False, und=dict(a0=0.3, lam=1.3, X0=-1.0, Lx=9.0), amp=0.5, boost=391.0),
    "env_m3": dict(setup="env_m3", ions=False, und=None, amp=0.5),      # envelope solver with +-1 modes
    # space-charge demo stage 1: quasi-static kick of a px = 50 beam (poiss_corr_stat, maxwell_init_push, field_drift)
    "static_m2": dict(setup="static_m2", ions=False, und=None, amp=0.0, boost=50.0),
    # FEL stage with its moving frame (doc/tests/fel-testrun.py:61-63): 'Staged' window every step -- half a shift in
    # frame_act(istep) before push_coords, half in frame_act(istep, 'stage2') between project_current and
    # project_density (chimera_main.py:40-51, 82-92, 292-304).  Beam close to the axis: the analytic undulator field
    # grows like cosh(ku y) (devices.f90:196-197).
    "env_m1_win": dict(setup="env_m1", ions=False, und=dict(a0=0.3, lam=1.3, X0=-1.0, Lx=9.0), amp=0.03, beam=True,
                       window=dict(Velocity=0.999, Staged=True)),
}

TABLES = ("In", "InCurr", "Out", "DpS2S", "DmS2S", "DepFact", "PoissFact", "kx", "kx_env", "Rgrid", "Xgrid", "VGrid")


def species_dict(cfg, **extra):
    d = {"Grid": cfg["Grid"], "TimeStep": cfg["TimeStep"]}
    if "Xchunked" in cfg:
        d["Xchunked"] = cfg["Xchunked"]
    d.update(extra)
    return d


def snapshot(prefix, run, out, grids=("J", "Rho", "BckGrndRho", "EB", "EG_fb")):
    sol = run.Solvers[0]
    for k in grids:
        if k in sol.Data:
            out["%s_%s" % (prefix, k)] = np.array(sol.Data[k], order="F")
    sp = run.Particles[0]
    for k in ("coords", "coords_halfstep", "momenta", "weights"):
        out["%s_%s" % (prefix, k)] = np.array(sp.Data[k], order="F")
    if hasattr(sp, "chunks"):
        out["%s_chunks" % prefix] = np.array(sp.chunks)


def generate(name, case, R, ofim):
    from util import SETUPS, plasma, seed_fields
    from chimera_b200.solver_setup import SolverSetup

    cfg = copy.deepcopy(SETUPS[case["setup"]])
    np.random.seed(20260101)
    solver = R.Solver(copy.deepcopy(cfg))
    out = {"cfg": np.array(json.dumps({"case": case, "nsteps": NSTEPS}))}
    for k in TABLES:
        out["tab_" + k] = np.array(solver.Args[k], order="F")
    out["tab_PSATD_E"] = np.array(solver.Data["PSATD_E"], order="F")
    out["tab_PSATD_G"] = np.array(solver.Data["PSATD_G"], order="F")
    for i, v in enumerate(solver.Args["DepProj"][1:]):
        out["tab_DepProj%d" % (i + 1)] = np.array(v)

    S = SolverSetup(copy.deepcopy(cfg))  # only used to shape the seeded inputs
    x, p, w = plasma(S, 2, 2, 11)
    p[0] += case.get("boost", 0.0)
    if case.get("beam"):
        a, rng, n = S.Args, np.random.default_rng(11), 3000
        xs = a["leftX"] + (0.25 + 0.5 * rng.random(n)) * (a["rightX"] - a["leftX"])
        r, th = 0.3 * np.sqrt(rng.random(n)), 2 * np.pi * rng.random(n)
        x = np.asfortranarray(np.vstack((xs, r * np.cos(th), r * np.sin(th))))
        p = np.asfortranarray(np.vstack((391.0 * (1 + 1e-4 * rng.standard_normal(n)), 2e-5 * 391 * rng.standard_normal(n),
                                         2e-5 * 391 * rng.standard_normal(n))))
        w = -1e-4 * (1 + 1e-3 * rng.random(n))
    eg0 = seed_fields(S, 12, case["amp"])
    out["in_coords"], out["in_momenta"], out["in_weights"], out["in_EG_fb"] = x, p, w, eg0
    solver.Data["EG_fb"][:] = eg0

    e_in = species_dict(cfg)
    if case["und"]:
        u = case["und"]
        e_in["Devices"] = ([ofim.undul_analytic, np.array([u["a0"], u["lam"], u["X0"], u["Lx"]])],)
    electrons = R.Specie(e_in)
    electrons.add_particles(x.copy(order="F"), p.copy(order="F"), w.copy())
    parts = [electrons]
    if case["ions"]:
        xi, pi_, wi = plasma(S, 2, 2, 18)
        out["in_ion_coords"], out["in_ion_weights"] = xi, -wi
        ions = R.Specie(species_dict(cfg, Charge=1.0, Mass=1886.0, Features=("Still",)))
        ions.add_particles(xi.copy(order="F"), 0 * pi_, -wi.copy())
        parts.append(ions)
    # a window that never moves: frame_act still runs every step (chimera_main.py:292-304) and, for a
    # SpaceCharge solver, re-deposits the background of the still species (postframe_corr :277-284)
    frame = {"Velocity": 0.0}
    if case.get("window"):
        wn = case["window"]
        frame = {"TimeStep": cfg["TimeStep"], "Steps": 1, "Velocity": wn["Velocity"],
                 "Features": ("Staged",) if wn["Staged"] else ()}
    run = R.ChimeraRun({"Solvers": (solver,), "Particles": tuple(parts), "MovingFrames": (frame,)})
    snapshot("h", run, out, grids=("EG_fb",))
    for i in range(1, NSTEPS + 1):
        run.make_step(i)
    snapshot("s", run, out)
    if case.get("window"):
        # the reference's own Diagnostics on the final state (moduls/diagnostics.py:109-207): field energy per kx,
        # power per x with the azimuthal 'Spot' profile (fb_vec_out + intens_profo), beam centroid / rms / emittance
        D = R.Diagnostics(run, (), out_folder=None)
        out["d_nrg"] = np.array(D.nrg_out({"Features": ("Return",)})[0])
        pwr, spot = D.pwr_out({"Features": ("Return", "Spot")})[0]
        out["d_pwr"], out["d_spot"] = np.array(pwr), np.array(spot)
        out["d_env"] = np.array(D.get_beam_envelops()[0])
    return out


# LPA stage with its moving window (doc/tests/lpa-testrun.py:41-64, shrunk): both species start EMPTY, the plasma
# enters through 'AddPlasma' as the window moves every LPA_WIND['Steps'] steps; absorbing layer on the left
LPA_NSTEPS = 13
LPA_WIND = dict(Steps=3, AbsorbLayer=8)
LPA_PROFILE = ([1.0, 1.3, 30.0], [0.0, 1.0, 1.0])  # np.interp nodes of the density profile along x


def generate_lpa(R, ofim):
    from util import SETUPS, seed_fields
    from chimera_b200.solver_setup import SolverSetup

    cfg = copy.deepcopy(SETUPS["real_m2"])
    np.random.seed(20260102)
    solver = R.Solver(copy.deepcopy(cfg))
    S = SolverSetup(copy.deepcopy(cfg))
    eg0 = seed_fields(S, 14, 0.5)
    solver.Data["EG_fb"][:] = eg0
    out = {"cfg": np.array(json.dumps({"case": dict(setup="real_m2", wind=LPA_WIND, profile=LPA_PROFILE), "nsteps": LPA_NSTEPS})),
           "in_EG_fb": eg0}
    e_in = species_dict(cfg, Density=0.005, FixedCell=(2, 2, 4), Features=("NoSorting",))
    i_in = species_dict(cfg, Density=0.005, FixedCell=(2, 2, 4), Charge=1, Mass=1886, Features=("NoSorting", "Still"))
    electrons, ions = R.Specie(e_in), R.Specie(i_in)
    adds = []  # (step, species index, coords, momenta, weights) of every add_particles call
    step = [0]
    for idx, sp in enumerate((electrons, ions)):
        orig = sp.add_particles

        def rec(coords, momenta, weights, idx=idx, orig=orig):
            adds.append((step[0], idx, np.array(coords, order="F"), np.array(momenta, order="F"), np.array(weights)))
            return orig(coords, momenta, weights)

        sp.add_particles = rec
    prof_x, prof_y = LPA_PROFILE
    wind = {"TimeStep": cfg["TimeStep"], "Steps": LPA_WIND["Steps"], "AbsorbLayer": LPA_WIND["AbsorbLayer"],
            "AddPlasma": lambda x: np.interp(x, prof_x, prof_y), "Features": ("IonsOnTop",)}
    run = R.ChimeraRun({"Solvers": (solver,), "Particles": (electrons, ions), "MovingFrames": (wind,)})
    out["h_EG_fb"] = np.array(solver.Data["EG_fb"], order="F")
    out["shiftX"] = np.array(wind["shiftX"])
    for i in range(1, LPA_NSTEPS + 1):
        step[0] = i
        run.make_step(i)
        if i in (7, LPA_NSTEPS):
            pre = "s%d" % i
            snapshot(pre, run, out)
            out[pre + "_ion_coords"] = np.array(ions.Data["coords"], order="F")
            out[pre + "_ion_weights"] = np.array(ions.Data["weights"])
    out["add_steps"] = np.array([a[0] for a in adds])
    out["add_species"] = np.array([a[1] for a in adds])
    for k, a in enumerate(adds):
        out["add%d_coords" % k], out["add%d_momenta" % k], out["add%d_weights" % k] = a[2], a[3], a[4]
    return out


def main():
    import ref_driver
    from oracle import fimera as ofim

    R = ref_driver.install(ofim)
    dst = os.path.join(ROOT, "tests", "golden")
    os.makedirs(dst, exist_ok=True)
    only = sys.argv[1:]
    if not only or "real_m2_lpa" in only:
        out = generate_lpa(R, ofim)
        path = os.path.join(dst, "real_m2_lpa.npz")
        np.savez_compressed(path, **out)
        print("%-8s %6.1f kB  adds at steps %s  electrons %d  |EG_fb| %.6e" % (
            "real_m2_lpa", os.path.getsize(path) / 1e3, sorted(set(out["add_steps"].tolist())),
            out["s%d_weights" % LPA_NSTEPS].size, np.linalg.norm(out["s%d_EG_fb" % LPA_NSTEPS].ravel())))
    for name, case in CASES.items():
        if only and name not in only:
            continue
        out = generate(name, case, R, ofim)
        path = os.path.join(dst, name + ".npz")
        np.savez_compressed(path, **out)
        print("%-8s %6.1f kB  particles %d -> %d  |EG_fb| %.6e" % (
            name, os.path.getsize(path) / 1e3, out["in_weights"].size, out["s_weights"].size,
            np.linalg.norm(out["s_EG_fb"].ravel())))


if __name__ == "__main__":
    main()
