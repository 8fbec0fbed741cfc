"""GPU parity AT THE SHAPES BASELINE.json NAMES (tests/baseline_shapes.py): the FEL stage (Nx=120, Nr=85 after the
R cut, envelope), the LPA stage (Nx=528, Nr=65, 2 modes), the space-charge drift (Nx=304, Nr=301, 2 modes, 1.4e6
particles, 'StaticKick' and 'SpaceCharge') and one step of the LWFA synthetic case (4096 x 513 x 3, 3.3e7
particles).  For each: every spectral entry point through the C ABI against the oracle (Nkr = 300 is ragged in
every GEMM tile size), one make_halfstep + make_step of the resident engine against the reference sequence on the
oracle at 1e-12, and 100 steps on the integrated diagnostics at 1e-6 (north_star)."""
import copy

import numpy as np
import pytest

import baseline_shapes as B
from pic_ref import RefRun, RefSpecies
from util import TOL, assert_close, carrier_tol, crandn, match, rel_l2
from chimera_b200.solver_setup import SolverSetup

pytestmark = pytest.mark.gpu

CASES = ["c1a", "c1b", "c2_static", "c2_pic"]


def case_def(case):
    return {"c1a": B.c1a_fel, "c1b": B.c1b_lpa, "c2_static": lambda: B.c2_space_charge("static"),
            "c2_pic": lambda: B.c2_space_charge("pic"), "c3": B.c3_lwfa}[case]()


def case_species(S, c, seed=7):
    if c["species"] is not None:
        return c["species"]
    if "beam" in c:
        return B.gaussian_beam(S, seed=seed, **c["beam"])
    p = c["plasma"]
    return B.fill_plasma(S, p["cell"], p["density"], seed, ions=p.get("ions", True))


def build_case(ofim, case, engine=True, sub=None):
    """(S, ref, eng, c): the reference sequence on the oracle and the resident engine on the same inputs"""
    from chimera_b200.engine import Engine

    c = case_def(case)
    S = SolverSetup(copy.deepcopy(c["cfg"]))
    species = case_species(S, c)
    if sub:  # entry-point tests: a subset of the particles is enough
        species = [dict(s, coords=s["coords"][:, ::sub], momenta=s["momenta"][:, ::sub], weights=s["weights"][::sub]) for s in species]
    eg0 = S.add_gauss_beam(ofim, c["laser"]) if c["laser"] else S.zeros_fb(6)
    dev = c["device"]
    ref_sp = [RefSpecies(s["coords"], s["momenta"], s["weights"], charge=s["charge"], mass=s["mass"], still=s["still"],
                         device=(getattr(ofim, dev[0]), dev[1]) if (dev and not s["still"]) else None) for s in species]
    ions = any(s["still"] for s in species)
    ref = RefRun(ofim, S, ref_sp, background=ions)
    ref.EG_fb[:] = eg0
    if c["window"]:
        v, staged = c["window"]
        dt = S.Args["dt"]
        ref.window = (0.5 * v * dt, 0.5 * v * dt) if staged else (v * dt, 0.0)
    eng = None
    if engine:
        und = dict(zip(("a0", "lambda", "X0", "Lx"), dev[1])) if dev else None
        eng = Engine(S, undulator=und)
        for s in species:
            eng.add_species(s["coords"], s["momenta"], s["weights"], charge=s["charge"], mass=s["mass"], still=s["still"])
        eng.upload("EG_fb", eg0)
        if c["window"]:
            eng.set_window(c["window"][0], staged=c["window"][1])
    return S, ref, eng, c


def compare(ref, eng, tol, names):
    tol = carrier_tol(ref.S, tol)
    table = {"J": ref.J, "Rho": ref.Rho, "EG_fb": ref.EG_fb, "J_fb": ref.J_fb, "EB": ref.EB}
    for n in names:
        t = 20 * tol if (ref.env and n in ("J", "J_fb")) else tol  # see test_gpu_engine.compare_state
        got = eng.download(n)
        if n == "Rho" and ref.background:
            # ions sit on top of the electrons ('IonsOnTop'): Rho = background + electrons is a remainder of two
            # cancelling deposits, so the error is measured against what was deposited
            err = np.linalg.norm((got - table[n]).ravel()) / np.linalg.norm(ref.Bck.ravel())
            assert err <= t, "Rho: error %.3e of |BckGrndRho| > %.1e" % (err, t)
            continue
        assert_close(got, table[n], t, n)
    x, xh, p, w = eng.particles(0)
    s = ref.sp[0]
    perm = match(s.weights, w)
    assert_close(p[:, perm], s.momenta, tol, "momenta")
    assert_close(x[:, perm], s.coords, tol, "coords")
    assert_close(xh[:, perm], s.coords_halfstep, tol, "coords_halfstep")


def both(ofim, gfim, name, *args):
    def cp(a):
        return [x.copy(order="F") if isinstance(x, np.ndarray) else x for x in a]
    return getattr(ofim, name)(*cp(args)), getattr(gfim, name)(*cp(args))


@pytest.mark.parametrize("case", ["c1a", "c1b", "c2_pic"])
def test_entry_points_at_baseline_shape(ofim, gfim, case):
    """fb_io / fb_math / maxwell_solvers / grid_deps entry points at the configuration's own (Nx, Nr, Nkr, M)"""
    c = case_def(case)
    S = SolverSetup(copy.deepcopy(c["cfg"]))
    a = S.Args
    sfx = "_env" if S.env else ""
    rng = np.random.default_rng(101)
    V, Sc = crandn(rng, S.shape_fb + (3,)), crandn(rng, S.shape_fb)
    for name, args in (("fb_rot" + sfx, (S.zeros_fb(3), V, *a["FBDiff"])), ("fb_grad" + sfx, (S.zeros_fb(3), Sc, *a["FBDiff"])),
                       ("fb_div" + sfx, (S.zeros_fb(), V, *a["FBDiff"])), ("fb_graddiv" + sfx, (V, *a["FBDiff"]))):
        ro, rg = both(ofim, gfim, name, *args)
        assert_close(rg, ro, what="%s @%s" % (name, case))
    J, R = crandn(rng, S.shape_sp + (3,)), crandn(rng, S.shape_sp)
    ro, rg = both(ofim, gfim, "fb_vec_in", S.zeros_fb(3), J, a["leftX"], *a["FBCurrIn"])
    assert_close(rg, ro, what="fb_vec_in @" + case)
    ro, rg = both(ofim, gfim, "fb_scl_in", S.zeros_fb(), R, a["leftX"], *a["FBIn"])
    assert_close(rg, ro, what="fb_scl_in @" + case)
    EG, Bf = crandn(rng, S.shape_fb + (6,)), crandn(rng, S.shape_fb + (3,))
    ro, rg = both(ofim, gfim, "fb_eb_out", S.zeros_sp(6), EG, Bf, a["leftX"], *a["FBout"])
    assert_close(rg, ro, what="fb_eb_out @" + case)
    ro, rg = both(ofim, gfim, "fb_vec_out", EG[:, :, :, 3:], a["leftX"], *a["FBout"])
    assert_close(rg, ro, what="fb_vec_out(slice) @" + case)
    g1, g2 = crandn(rng, S.shape_fb + (3,)), crandn(rng, S.shape_fb + (3,))
    if S.space_charge:
        ro, rg = both(ofim, gfim, "maxwell_push_with_spchrg", EG, V, g1, g2, S.PSATD_E, S.PSATD_G)
    else:
        ro, rg = both(ofim, gfim, "maxwell_push_wo_spchrg", EG, V, S.PSATD_E, S.PSATD_G)
    assert_close(rg, ro, what="maxwell_push @" + case)
    ro, rg = both(ofim, gfim, "poiss_corr", V, g1, g2, EG[..., :3], a["dt_inv"], a["PoissFact"])
    assert_close(rg, ro, what="poiss_corr @" + case)
    # particle entry points on the configuration's own particles (every 4th), chunk-sorted as the driver does
    from util import chunk_sorted

    sp = case_species(S, c)[0]
    x, p, w = sp["coords"][:, ::4], sp["momenta"][:, ::4], sp["weights"][::4]
    nchnk, guards = a["Xchunked"]
    x, p, w, chunks = chunk_sorted(S, np.asfortranarray(x), np.asfortranarray(p), np.ascontiguousarray(w), ofim, nchnk)
    tol = carrier_tol(S)
    ro, rg = both(ofim, gfim, "dep_curr" + sfx + "_chnk", x, p, w, S.zeros_sp(3), chunks, guards, a["leftX"], *a["DepProj"])
    assert_close(rg, ro, 20 * tol if S.env else tol, what="dep_curr_chnk @" + case)
    ro, rg = both(ofim, gfim, "dep_dens" + sfx + "_chnk", x, w, S.zeros_sp(), chunks, guards, a["leftX"], *a["DepProj"])
    assert_close(rg, ro, 20 * tol if S.env else tol, what="dep_dens_chnk @" + case)
    F = crandn(rng, S.shape_sp + (6,))
    ro, rg = both(ofim, gfim, "proj_fld" + sfx, x, w, F, np.zeros((6, w.size), order="F"), a["leftX"], *a["DepProj"])
    assert_close(rg, ro, tol, what="proj_fld @" + case)


@pytest.mark.parametrize("case", CASES)
def test_one_step_at_baseline_shape(ofim, gfim, case):
    """make_halfstep + make_step (chimera_main.py:61-92) on the full configuration, fields and particles at 1e-12"""
    S, ref, eng, c = build_case(ofim, case)
    names = ("J", "Rho", "EG_fb", "EB") if (ref.space_charge or ref.static_kick) else ("J", "EG_fb", "EB")
    ref.make_halfstep(px0=c["px0"])
    eng.make_halfstep(px0=c["px0"], background=ref.background)
    compare(ref, eng, TOL, names)
    ref.make_step()
    eng.step(1)
    compare(ref, eng, 2 * TOL, names)
    eng.close()


def diagnostics(S, eg_fb, eb_axis, x, p, w):
    """total charge, field energy (nrg_out, diagnostics.py:109-124), on-axis wake amplitude (|Re EB[:,0,0,0]|
    summed), total particle energy, mean / rms of gamma and a 12-window energy spectrum (Gaussian windows: a hard
    bin edge turns a 1e-13 difference of one particle into a 1/N jump)"""
    gam = np.sqrt(1 + (p ** 2).sum(0))
    g0 = (w * gam).sum() / w.sum()
    sg = np.sqrt(max((w * (gam - g0) ** 2).sum() / w.sum(), 0.0)) + 1e-3 * abs(g0 - 1.0) + 1e-12
    centres = g0 + sg * np.linspace(-3.0, 3.0, 12)
    spec = (w[None, :] * np.exp(-((gam[None, :] - centres[:, None]) / (0.7 * sg)) ** 2)).sum(1)
    nrg = (np.abs(eg_fb[..., :3]) ** 2 * S.Args["EnergyFact"][..., None]).sum()
    wake = np.abs(eb_axis.real).sum()
    return np.concatenate(([w.sum(), nrg, wake, (w * gam).sum(), g0], spec)), (g0, sg)


def run_100(ref, eng, c, nsteps):
    """`nsteps` steps on both; the LPA case's window acts every frame['Steps'] steps (chimera_main.py:250-304) and
    feeds fresh plasma into the cells that entered on the right"""
    frame = (c.get("plasma") or {}).get("frame")
    if not frame:
        for _ in range(nsteps):
            ref.make_step()
        eng.step(nsteps)
        return
    S, a = ref.S, ref.a
    every, done, k = frame["Steps"], 0, 0
    wind = {"shiftX": every * a["dx"] * 1.0, "AbsorbLayer": frame["AbsorbLayer"], "Features": ("IonsOnTop",)}
    while done < nsteps:
        n = min(every - 1 if done == 0 else every, nsteps - done)
        for _ in range(n):
            ref.make_step()
        eng.step(n)
        done += n
        if done >= nsteps:
            break
        k += 1
        shifted = copy.copy(S)  # the fresh cells are laid out on the grid as it will be after the move
        shifted.Args = dict(a, leftX=a["leftX"] + wind["shiftX"])
        new = B.fill_plasma(shifted, c["plasma"]["cell"], c["plasma"]["density"], 1000 + k, ix0=a["Nx"] - 1 - every, ix1=a["Nx"] - 1)
        add = {i: (s["coords"], s["momenta"], s["weights"]) for i, s in enumerate(new)}
        ref.frame_act(wind, add)
        eng.frame_act(wind, add, background=ref.background)


@pytest.mark.parametrize("case,nsteps", [("c1a", 100), ("c1b", 100), ("c2_static", 21), ("c2_pic", 100)])
def test_100_steps_at_baseline_shape(ofim, gfim, case, nsteps):
    """integrated diagnostics within 1e-6 after 100 steps (the static-kick stage of the demo has 21 steps in all)"""
    S, ref, eng, c = build_case(ofim, case)
    ref.make_halfstep(px0=c["px0"])
    eng.make_halfstep(px0=c["px0"], background=ref.background)
    run_100(ref, eng, c, nsteps)
    x, xh, p, w = eng.particles(0)
    s = ref.sp[0]
    assert w.size == s.weights.size
    d_ref, _ = diagnostics(S, ref.EG_fb, ref.EB[:, 0, 0, 0], s.coords, s.momenta, s.weights)
    d_eng, _ = diagnostics(S, eng.download("EG_fb"), eng.lineout("EB", 0, 0, 0), x, p, w)
    scale = np.maximum(np.abs(d_ref), 1e-3 * np.abs(d_ref[5:]).max() * (np.arange(d_ref.size) >= 5))
    scale = np.where(scale > 0, scale, 1.0)
    err = np.abs(d_eng - d_ref) / scale
    assert err.max() < 1e-6, (case, err, d_eng, d_ref)
    assert rel_l2(eng.download("EG_fb"), ref.EG_fb) < 1e-6
    perm = match(s.weights, w)
    assert rel_l2(p[:, perm], s.momenta) < 1e-6
    eng.close()


def test_c3_lwfa_one_step(ofim, gfim):
    """BASELINE configs[2] at full grid size, 16 per cell (3.3e7 macro-particles): make_halfstep + one make_step of
    the resident engine against the reference sequence on the oracle, 1e-12 on fields and momenta"""
    S, ref, eng, c = build_case(ofim, "c3")
    names = ("J", "Rho", "EG_fb", "EB")
    ref.make_halfstep()
    eng.make_halfstep()
    compare(ref, eng, TOL, names)
    ref.make_step()
    eng.step(1)
    compare(ref, eng, 2 * TOL, names)
    eng.close()
