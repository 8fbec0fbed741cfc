"""GPU parity of the device-resident engine (csrc/engine.cu) against the reference step sequence
(tests/pic_ref.py, a restatement of ChimeraRun.make_halfstep/make_step) run on the CPU oracle.

Tolerances (BASELINE.json north_star): fields and particle momenta after one step within 1e-12
relative L2; integrated diagnostics after 100 steps within 1e-6 (atomic ordering is the only source
of divergence)."""
import copy

import numpy as np
import pytest

from pic_ref import RefRun, RefSpecies
from util import SETUPS, TOL, assert_close, carrier_tol, match, plasma, rel_l2, seed_fields
from chimera_b200.solver_setup import SolverSetup

pytestmark = pytest.mark.gpu


def build_pair(ofim, name, seed, ppc=(2, 2), still_ions=False, undulator=None, amp=0.5):
    from chimera_b200.engine import Engine

    S = SolverSetup(copy.deepcopy(SETUPS[name]))
    x, p, w = plasma(S, ppc[0], ppc[1], seed)
    if name == "env_m1":
        p[0] += 391.0  # the FEL beam; slow particles turn this setup into a round-off amplifier (tools/gen_golden.py)
    eg0 = seed_fields(S, seed + 1, amp)
    dev = None
    if undulator:
        dev = (ofim.undul_analytic, [undulator[k] for k in ("a0", "lambda", "X0", "Lx")])
    sp = [RefSpecies(x, p, w, device=dev)]
    eng = Engine(S, undulator=undulator)
    eng.add_species(x, p, w)
    if still_ions:  # an independent sample, so that the net charge density is not identically zero
        xi, pi_, wi = plasma(S, ppc[0], ppc[1], seed + 7)
        sp.append(RefSpecies(xi, 0 * pi_, -wi, charge=1.0, mass=1886.0, still=True))
        eng.add_species(xi, 0 * pi_, -wi, charge=1.0, mass=1886.0, still=True)
    ref = RefRun(ofim, S, sp, background=still_ions)
    ref.EG_fb[:] = eg0
    eng.upload("EG_fb", eg0)
    return S, ref, eng


def compare_state(ref, eng, tol, names=("J", "Rho", "EG_fb", "J_fb", "EB")):
    tol = carrier_tol(ref.S, tol)
    for n in names:
        want = {"J": ref.J, "Rho": ref.Rho, "EG_fb": ref.EG_fb, "J_fb": ref.J_fb, "EB": ref.EB, "B_fb": ref.B_fb}[n]
        if n == "Rho" and not ref.space_charge:
            continue
        # the envelope deposit sums particle terms that carry exp(-i kx0 x) with kx0 x ~ 1e3..1e6 rad: the
        # deposited J is a small remainder of cancelling terms, so its relative error is the summation-order
        # error (1e-16 * sum|terms|) amplified by the cancellation; fields and momenta stay at `tol`
        t = 20 * tol if (ref.env and n in ("J", "J_fb")) else tol
        assert_close(eng.download(n), want, t, n)
    x, xh, p, w = eng.particles(0)
    s = ref.sp[0]
    assert w.shape == s.weights.shape
    perm = match(s.weights, w)
    assert_close(p[:, perm], s.momenta, tol, "momenta")
    assert_close(x[:, perm], s.coords, tol, "coords")
    assert_close(xh[:, perm], s.coords_halfstep, tol, "coords_halfstep")


@pytest.mark.parametrize("name,ions", [("real_m2", True), ("real_m3", False), ("env_m3", False), ("env_m1", False)])
def test_engine_halfstep_and_one_step(ofim, gfim, name, ions):
    und = dict(a0=0.3, **{"lambda": 1.3}, X0=-1.0, Lx=9.0) if name == "env_m1" else None
    S, ref, eng = build_pair(ofim, name, 11, still_ions=ions, undulator=und)
    ref.make_halfstep()
    eng.make_halfstep(background=ions)
    compare_state(ref, eng, TOL)
    if ref.chunked:
        assert np.array_equal(eng.chunks(0), ref.sp[0].chunks)
    ref.make_step()
    eng.step(1)
    compare_state(ref, eng, TOL)
    eng.close()


def diagnostics(S, eg_fb, x, p, w, gam_ref):
    """integrated diagnostics: total charge, field energy (diagnostics.py:109 nrg_out), wake amplitude
    proxy (mode-0 Ex spectral amplitude), total particle energy and a 12-point energy spectrum.  The
    spectrum uses Gaussian windows instead of hard histogram bins: with ~2000 test particles a single
    particle crossing a bin edge would change a bin by 5e-4, which says nothing about parity."""
    gam = np.sqrt(1 + (p ** 2).sum(0))
    centres = 1.0 + (gam_ref - 1.0) * np.linspace(0.0, 3.0, 12)
    width = 0.5 * (gam_ref - 1.0)
    spec = (w[None, :] * np.exp(-((gam[None, :] - centres[:, None]) / width) ** 2)).sum(1)
    nrg = (np.abs(eg_fb[..., :3]) ** 2 * S.Args["EnergyFact"][..., None]).sum()
    wake = np.abs(eg_fb[:, :, 0, 0]).sum()
    return np.concatenate(([w.sum(), nrg, wake, (w * gam).sum()], spec))


@pytest.mark.parametrize("name", ["real_m2", "env_m3"])
def test_engine_100_steps_diagnostics(ofim, gfim, name):
    # moderate field amplitude: at a ~ 1 the single-particle orbits of this tiny, noisy test plasma are
    # chaotic and round-off differences grow by orders of magnitude in 100 steps on ANY two machines
    S, ref, eng = build_pair(ofim, name, 21, still_ions=(name == "real_m2"), amp=0.03)
    ref.make_halfstep()
    eng.make_halfstep(background=(name == "real_m2"))
    for _ in range(100):
        ref.make_step()
    eng.step(100)
    x, xh, p, w = eng.particles(0)
    gam_ref = float(np.sqrt(1 + (ref.sp[0].momenta ** 2).sum(0)).mean()) + 1e-3
    d_eng = diagnostics(S, eng.download("EG_fb"), x, p, w, gam_ref)
    d_ref = diagnostics(S, ref.EG_fb, ref.sp[0].coords, ref.sp[0].momenta, ref.sp[0].weights, gam_ref)
    scale = np.maximum(np.abs(d_ref), 1e-3 * np.abs(d_ref).max())
    err = np.abs(d_eng - d_ref) / scale
    assert err.max() < 1e-6, (err, d_eng, d_ref)
    assert rel_l2(eng.download("EG_fb"), ref.EG_fb) < 1e-6
    eng.close()


def test_engine_100_steps_fel_stage(ofim, gfim):
    """BASELINE configs[0]/[3] in miniature: envelope solver, gamma = 391 beam close to the axis (the analytic
    undulator field grows like cosh(ku y), devices.f90:196-197), 'Staged' window moving with the beam every step
    (doc/tests/fel-testrun.py:61-63), 100 steps resident on the device against the reference sequence on the oracle;
    integrated diagnostics (total charge, field energy = nrg_out, on-axis amplitude, beam energy and its spectrum)
    within 1e-6."""
    from chimera_b200.engine import Engine

    S = SolverSetup(copy.deepcopy(SETUPS["env_m1"]))
    a = S.Args
    rng = np.random.default_rng(25)
    n = 3000
    x = a["leftX"] + (0.25 + 0.5 * rng.random(n)) * (a["rightX"] - a["leftX"])
    r, th = 0.3 * np.sqrt(rng.random(n)), 2 * np.pi * rng.random(n)
    coords = np.asfortranarray(np.vstack((x, r * np.cos(th), r * np.sin(th))))
    mom = np.asfortranarray(np.vstack((391.0 * (1 + 1e-4 * rng.standard_normal(n)), 2e-5 * 391 * rng.standard_normal(n),
                                       2e-5 * 391 * rng.standard_normal(n))))
    w = -1e-4 * (1 + 1e-3 * rng.random(n))
    und = dict(a0=0.3, **{"lambda": 1.3}, X0=-1.0, Lx=9.0)
    ref = RefRun(ofim, S, [RefSpecies(coords, mom, w, device=(ofim.undul_analytic, [0.3, 1.3, -1.0, 9.0]))])
    eng = Engine(S, undulator=und)
    eng.add_species(coords, mom, w)
    eg0 = seed_fields(S, 26, 0.03)
    ref.EG_fb[:] = eg0
    eng.upload("EG_fb", eg0)
    v, dt = 0.999, a["dt"]
    ref.window = (0.5 * v * dt, 0.5 * v * dt)
    eng.set_window(v, staged=True)
    ref.make_halfstep()
    eng.make_halfstep()
    for _ in range(100):
        ref.make_step()
    eng.step(100)
    xe, xh, pe, we = eng.particles(0)
    assert we.size == ref.sp[0].weights.size == n
    gam_ref = float(np.sqrt(1 + (ref.sp[0].momenta ** 2).sum(0)).mean()) * (1 + 1e-4)
    d_eng = diagnostics(S, eng.download("EG_fb"), xe, pe, we, gam_ref)
    d_ref = diagnostics(S, ref.EG_fb, ref.sp[0].coords, ref.sp[0].momenta, ref.sp[0].weights, gam_ref)
    scale = np.maximum(np.abs(d_ref), 1e-3 * np.abs(d_ref).max())
    err = np.abs(d_eng - d_ref) / scale
    assert err.max() < 1e-6, (err, d_eng, d_ref)
    assert rel_l2(eng.download("EG_fb"), ref.EG_fb) < 1e-6
    perm = match(ref.sp[0].weights, we)
    assert_close(pe[:, perm], ref.sp[0].momenta, 1e-6, "momenta after 100 steps")
    eng.close()


def test_engine_dropin_sequence_matches(ofim, gfim):
    """the host-buffer drop-in (chimera_b200.fimera) driven by the same step sequence"""
    from chimera_b200.engine import Engine  # noqa: F401

    S = SolverSetup(copy.deepcopy(SETUPS["real_m2"]))
    x, p, w = plasma(S, 2, 2, 31)
    eg0 = seed_fields(S, 32)
    runs = []
    for fim in (ofim, gfim):
        r = RefRun(fim, S, [RefSpecies(x, p, w)])
        r.EG_fb[:] = eg0
        r.make_halfstep()
        r.make_step()
        runs.append(r)
    assert_close(runs[1].EG_fb, runs[0].EG_fb, TOL, "EG_fb")
    assert_close(runs[1].EB, runs[0].EB, TOL, "EB")
    perm = match(runs[0].sp[0].weights, runs[1].sp[0].weights)
    assert_close(runs[1].sp[0].momenta[:, perm], runs[0].sp[0].momenta, TOL, "momenta")


@pytest.mark.parametrize("name,ions,want_half", [("real_m2", True, True), ("real_m3", False, True), ("env_m1", False, True),
                                                 ("real_m2", True, False)])
def test_engine_step_host_matches_reference_sequence(ofim, gfim, name, ions, want_half):
    """chimera_engine_step_host: the whole PIC state crosses PCIe every step (host numpy arrays in and
    out, copies pipelined with the kernels); 5 steps include a re-binning step for Xchunked=(4,3).
    want_half = False: coords_halfstep not requested back (NULL), everything else unchanged."""
    und = dict(a0=0.3, **{"lambda": 1.3}, X0=-1.0, Lx=9.0) if name == "env_m1" else None
    S, ref, eng = build_pair(ofim, name, 51, still_ions=ions, undulator=und)
    ref.make_halfstep()
    eng.make_halfstep(background=ions)
    x, xh, p, w = eng.particles(0)
    eg = eng.download("EG_fb")
    g = eng.download("gradRho_fb_nxt") if eng.cfg.space_charge else None
    eng.pin(x, xh, p, w, eg, g)
    n = x.shape[1]
    for _ in range(5):
        ref.make_step()
        n = eng.step_host(x[:, :n], xh[:, :n] if want_half else None, p[:, :n], w[:n], eg, g)  # leading columns stay Fortran-contiguous
    tol = carrier_tol(S, 5 * TOL)
    s = ref.sp[0]
    assert n == s.weights.shape[0]
    perm = match(s.weights, w[:n])
    assert_close(p[:, :n][:, perm], s.momenta, tol, "momenta")
    assert_close(x[:, :n][:, perm], s.coords, tol, "coords")
    if want_half:
        assert_close(xh[:, :n][:, perm], s.coords_halfstep, tol, "coords_halfstep")
    else:  # still on the device
        assert_close(eng.particles(0)[1][:, perm], s.coords_halfstep, tol, "coords_halfstep (engine)")
    assert_close(eg, ref.EG_fb, tol, "EG_fb")
    if g is not None:
        assert_close(g, ref.g_nxt, tol, "gradRho_fb_nxt")
    assert_close(eng.download("EB"), ref.EB, tol, "EB")
    eng.close()


def _exchange_slabs(engines, split=False):
    """what the NCCL all-gather does across ranks, emulated between engines that share one GPU; split: the E and B
    halves gathered separately (EB_gath = [half][rank][...]), as the pipelined multi-rank path does"""
    import torch

    torch.cuda.synchronize()
    for e in engines:
        e.sync()
    slabs = [e.device_tensor("EB_slab") for e in engines]
    if split:
        h = slabs[0].numel() // 2
        gath = torch.cat([s[:h] for s in slabs] + [s[h:] for s in slabs])
    else:
        gath = torch.cat(slabs)
    for e in engines:
        e.device_tensor("EB_gath").copy_(gath)
    torch.cuda.synchronize()


def _fields_out_slabs(engines, split):
    if not split:
        for e in engines:
            e.run("fields_out_a")
        _exchange_slabs(engines)
        for e in engines:
            e.run("fields_out_b")
        return
    for e in engines:
        e.run("fields_out_a", 1.0)
        e.run("fields_out_a", 2.0)
    _exchange_slabs(engines, split=True)
    for e in engines:
        e.run("fields_out_b", 2.0)  # any order
        e.run("fields_out_b", 1.0)


@pytest.mark.parametrize("name,world", [("real_m2", 2), ("real_m2", 4), ("env_m3", 2)])
def test_damp_field_on_kx_slab_engines(gfim, name, world):
    """Solver.damp_field (fb_filtr: an x-space window) on engines that hold kx slabs: the slabs are gathered, the full
    rows filtered and the own rows kept (chimera_engine_damp_field_slab); against the unsharded engine, after a window
    move so that the phase exp(i kx leftX) is not the initial one.  The all-gather is emulated on one GPU."""
    import torch

    from chimera_b200 import sharding
    from chimera_b200.engine import Engine

    S = SolverSetup(copy.deepcopy(SETUPS[name]))
    eg0 = seed_fields(S, 64)
    prof = S.get_damp_profile(8)
    full = Engine(S, slab=False)
    full.upload("EG_fb", eg0)
    full.move_window(0.15)
    full.damp_field(prof)
    want = full.download("EG_fb")
    full.close()
    assert rel_l2(want, eg0) > 1e-3  # the window did something
    engines = []
    for r in range(world):
        e = Engine(S, slab=(r, world))
        e.upload("EG_fb", eg0)
        e.move_window(0.15)
        e.damp_field_gather()  # allocates EG_gath / EG_full, uploads the full kx; no collective in a single process
        engines.append(e)
    torch.cuda.synchronize()
    for e in engines:
        e.sync()
    gath = torch.cat([e.device_tensor("EG_fb") for e in engines])
    for e in engines:
        e.device_tensor("EG_gath").copy_(gath)
    torch.cuda.synchronize()
    for r, e in enumerate(engines):
        e.damp_field(prof)
        rows = sharding.kx_slab_rows(S.Args["Nx"], r, world)
        assert_close(e.download("EG_fb"), want[rows], TOL, "EG_fb slab %d after damp_field" % r)
        e.close()


@pytest.mark.parametrize("name,world,split", [("real_m2", 2, False), ("real_m2", 4, True), ("real_m3", 5, False),
                                              ("env_m3", 2, True), ("real_m3", 5, True)])
def test_kx_slab_sharded_solve_matches_reference(ofim, gfim, name, world, split):
    """The spectral solve sharded by kx slab (mirror pairs of rows; x-FFT first, DHT / Poisson / PSATD / rot /
    backward DHT on the slab, all-gather, inverse x-FFT): `world` slab engines on one GPU, each with the full
    particle set, the all-gather emulated by device copies.  Halfstep + 2 steps against the oracle sequence."""
    from chimera_b200 import sharding
    from chimera_b200.engine import Engine

    S = SolverSetup(copy.deepcopy(SETUPS[name]))
    x, p, w = plasma(S, 2, 2, 61)
    eg0 = seed_fields(S, 62)
    ref = RefRun(ofim, S, [RefSpecies(x, p, w)])
    ref.EG_fb[:] = eg0
    engines = []
    for r in range(world):
        e = Engine(S, slab=(r, world))
        assert e.slab and e.cfg.nx_slab == S.Args["Nx"] // world
        e.add_species(x, p, w)
        e.upload("EG_fb", eg0)
        engines.append(e)

    def deposit_and_transform(e):
        e.run("deposit_J")
        if e.cfg.space_charge:
            e.run("deposit_rho", 1.0)
        e.run("fb_in_J")
        if e.cfg.space_charge:
            e.run("fb_in_rho")

    ref.make_halfstep()
    for e in engines:
        e.run("sort", 0.0)
        deposit_and_transform(e)
        if e.cfg.space_charge:
            c1, c2 = S.static_coeffs(0.0)
            e.upload("CPSATD1", c1)
            e.upload("CPSATD2", c2)
            e.run("init_push")
    _fields_out_slabs(engines, split)
    for e in engines:
        e.run("gather_push", 0.5)
    for istep in (1, 2):
        ref.make_step()
        for e in engines:
            e.run("push_coords")
            if e.cfg.sort_every > 0 and istep % e.cfg.sort_every == 0:
                e.run("sort", 1.0)
            deposit_and_transform(e)
            e.run("poisson")
            e.run("maxwell")
        _fields_out_slabs(engines, split)
        for e in engines:
            e.run("gather_push", 1.0)
    tol = carrier_tol(S, 3 * TOL)
    for r, e in enumerate(engines):
        rows = sharding.kx_slab_rows(S.Args["Nx"], r, world)
        assert_close(e.download("EG_fb"), ref.EG_fb[rows], tol, "EG_fb slab %d" % r)
        assert_close(e.download("J_fb"), ref.J_fb[rows], 20 * tol if S.env else tol, "J_fb slab %d" % r)
        assert_close(e.download("EB"), ref.EB, tol, "EB on rank %d" % r)
        xe, xhe, pe, we = e.particles(0)
        perm = match(ref.sp[0].weights, we)
        assert_close(pe[:, perm], ref.sp[0].momenta, tol, "momenta on rank %d" % r)
        e.close()


@pytest.mark.parametrize("name,ions", [("real_m2", True), ("real_m3", False), ("env_m3", False), ("env_m1", False)])
@pytest.mark.parametrize("nsteps", [3, 6])
def test_fused_particle_kernel_matches_reference_sequence(ofim, gfim, name, ions, nsteps):
    """Engine.step(n>1) runs gather + push_velocs of step k and push_coords + deposits of step k+1 as one kernel
    (csrc/particles_fused.cu); 6 steps cross a re-binning step for Xchunked=(4,3).  Also checked against the
    unfused engine, which must agree to summation order."""
    und = dict(a0=0.3, **{"lambda": 1.3}, X0=-1.0, Lx=9.0) if name == "env_m1" else None
    S, ref, eng = build_pair(ofim, name, 81, still_ions=ions, undulator=und)
    _, _, eng2 = build_pair(ofim, name, 81, still_ions=ions, undulator=und)
    eng2.set_fuse(False)
    ref.make_halfstep()
    for e in (eng, eng2):
        e.make_halfstep(background=ions)
    for _ in range(nsteps):
        ref.make_step()
    eng.step(nsteps)
    eng2.step(nsteps)
    compare_state(ref, eng, nsteps * TOL)
    for n in ("J", "EB", "EG_fb"):
        assert_close(eng.download(n), eng2.download(n), carrier_tol(S, 20 * nsteps * TOL if S.env else nsteps * TOL), n + " fused vs unfused")
    eng.close()
    eng2.close()


@pytest.mark.parametrize("name,staged", [("env_m1", True), ("env_m1", False), ("real_m2", True)])
@pytest.mark.parametrize("fuse", [True, False])
def test_engine_window_moving_every_step(ofim, gfim, name, staged, fuse):
    """MovingFrame with 'Steps': 1 (the FEL runs, doc/tests/fel-testrun.py:61-63): the grid origin advances inside
    every make_step -- half before push_coords and half between dep_curr and dep_dens for a 'Staged' frame
    (chimera_main.py:40-51, 83-87), so J, rho and the gathered fields live on three different window positions.
    The engine handles it inside step() (the fused kernel deposits on the moved grids); 8 steps cross a re-binning
    step.  The window velocity is not a multiple of dx/dt, so cell indices do change relative to the binning."""
    und = dict(a0=0.3, **{"lambda": 1.3}, X0=-1.0, Lx=9.0) if name == "env_m1" else None
    S, ref, eng = build_pair(ofim, name, 83, undulator=und, still_ions=False)
    v = 0.37 if name == "real_m2" else 0.999
    dt = S.Args["dt"]
    ref.window = (0.5 * v * dt, 0.5 * v * dt) if staged else (v * dt, 0.0)
    eng.set_window(v, staged=staged)
    eng.set_fuse(fuse)
    ref.make_halfstep()
    eng.make_halfstep()
    nsteps = 8
    for _ in range(nsteps):
        ref.make_step()
    eng.step(3)
    eng.step(nsteps - 3)
    compare_state(ref, eng, nsteps * TOL)
    eng.set_window(0.0)
    ref.window = (0.0, 0.0)
    ref.make_step()
    eng.step(1)
    compare_state(ref, eng, (nsteps + 1) * TOL)
    eng.close()


@pytest.mark.parametrize("name", ["real_m2", "env_m1"])
def test_engine_device_list(ofim, gfim, name):
    """NEXT-1: every kind of external device (devices.f90) between gather and push inside the resident engine,
    time-dependent ones included (t = i_step * TimeStep, species.py:274-277), through the separate kernels
    (first step, re-binning steps) and the fused kernel."""
    from chimera_b200.engine import Engine

    S = SolverSetup(copy.deepcopy(SETUPS[name]))
    a = S.Args
    x, p, w = plasma(S, 2, 2, 31)
    if name == "env_m1":
        p[0] += 391.0
    L = a["rightX"] - a["leftX"]
    nxm = 64
    dxm = L / (nxm - 8)
    amap = np.asfortranarray(np.vstack((np.sin(0.5 * np.arange(nxm)), np.cos(0.5 * np.arange(nxm)))) * 0.7)
    devs = [
        ("planewave", (np.array([0.4, 0.3 * L, a["leftX"] + 0.1 * L, 0.8 * L, 0.1 * L, 0.2, 0.5]),), {}),
        ("gaussbeam", (0.5, np.array([0.25 * L, 1.0, a["leftX"] + 0.4 * L, 0.1, -0.2, 0.2 * L, 1.0, 1.5])), {}),
        ("undul_mapped_tap", (amap, np.array([0.2 * L, a["leftX"] - 3 * dxm, dxm, 0.7 * L, 0.1])), {}),
        ("undul_analytic_taper", (np.array([0.9, 0.15 * L, a["leftX"] + 0.05 * L, 0.9 * L, -0.1]),), {}),
    ]
    ref = RefRun(ofim, S, [RefSpecies(x, p, w, devices=[(getattr(ofim, n),) + args for n, args, _ in devs])])
    eng = Engine(S)
    eng.add_species(x, p, w)
    for n, args, _ in devs:
        if n == "gaussbeam":
            eng.add_device(n, args[1], a0=args[0])
        elif n.startswith("undul_mapped"):
            eng.add_device(n, args[1], a0_map=args[0])
        else:
            eng.add_device(n, args[0])
    eg0 = seed_fields(S, 32, 0.3)
    ref.EG_fb[:] = eg0
    eng.upload("EG_fb", eg0)
    ref.make_halfstep()
    eng.make_halfstep()
    compare_state(ref, eng, TOL, names=("EB",))
    for _ in range(5):
        ref.make_step()
    eng.step(5)
    compare_state(ref, eng, 10 * TOL)
    eng.close()


@pytest.mark.parametrize("name,ions", [("real_m2", True), ("real_m3", False)])
def test_engine_moving_window(ofim, gfim, name, ions):
    """NEXT-2: the moving window on the device (chimera_main.py:250-304 frame_act stage 1): absorbing-layer field
    damping, grid shift, particle injection at the right edge, cull + re-binning, background / density redo -- twice,
    with steps in between, against the same sequence on the oracle."""
    S, ref, eng = build_pair(ofim, name, 61, still_ions=ions)
    a = S.Args
    ref.make_halfstep()
    eng.make_halfstep(background=ions)
    rng = np.random.default_rng(62)
    wind = {"shiftX": 4 * a["dx"], "AbsorbLayer": 24, "Features": ()}
    for rnd in range(2):
        for _ in range(3):
            ref.make_step()
        eng.step(3)
        # a layer of fresh plasma in the cells that enter on the right (gen_parts(Xsteps=...), species.py:171)
        n = 400
        xr = a["Xgrid"][-1] + wind["shiftX"]
        x = np.asfortranarray(np.vstack((xr - rng.random(n) * wind["shiftX"], (rng.random((2, n)) - 0.5) * 1.2 * a["Rgrid"].max())))
        p = np.asfortranarray(rng.standard_normal((3, n)) * 0.05)
        w = -np.abs(rng.random(n)) * 1e-3
        add = {0: (x, p, w)}
        if ions:
            add[1] = (x.copy(order="F"), np.zeros_like(p), -w)
        ref.frame_act(wind, add)
        eng.frame_act(wind, add, background=ions)
        assert eng.count(0) == ref.sp[0].weights.shape[0]
        assert_close(eng.download("EG_fb"), ref.EG_fb, 10 * TOL, "EG_fb after damp_field")
        assert_close(eng.download("Rho"), ref.Rho, 10 * TOL, "Rho after postframe_corr")
        if ions:
            assert_close(eng.download("BckGrndRho"), ref.Bck, 10 * TOL, "BckGrndRho")
    for _ in range(2):
        ref.make_step()
    eng.step(2)
    compare_state(ref, eng, 20 * TOL)
    eng.close()


@pytest.mark.parametrize("name", ["real_m2", "env_m3"])
def test_engine_device_diagnostics(ofim, gfim, name):
    """NEXT-3: the reductions behind nrg_out / get_beam_envelops / energy spectrum / on-axis line-outs computed on
    the device against the reference's numpy expressions (moduls/diagnostics.py:109-124, 174-207) on downloaded state"""
    S, ref, eng = build_pair(ofim, name, 71, amp=0.2)
    a = S.Args
    eng.make_halfstep()
    eng.step(5)
    eg = eng.download("EG_fb")
    dat = ((np.abs(eg[:, :, :, :3]) ** 2).sum(-1) * a["EnergyFact"]).sum(-1).sum(-1)
    want = np.r_[dat[dat.shape[0] // 2 + 1:], dat[:dat.shape[0] // 2 + 1]]
    for _ in range(2):  # second call: cached table
        assert_close(eng.nrg_out(), want, 1e-12, "nrg_out")
    _, xh, p, w = eng.particles(0)
    sw = w.sum()
    x0 = [(xh[c] * w).sum() / sw for c in range(3)]
    rms = [np.sqrt((xh[c] ** 2 * w).sum() / sw - x0[c] ** 2) for c in range(3)]
    emit = [np.sqrt((xh[c] ** 2 * w).sum() / sw * (p[c] ** 2 * w).sum() / sw - (xh[c] * p[c] * w).sum() ** 2 / sw ** 2) for c in range(3)]
    got = eng.get_beam_envelops(0)
    assert np.allclose(got, np.array([x0, rms, emit]), rtol=1e-9, atol=1e-12), (got, x0, rms, emit)
    assert np.isclose(eng.beam_moments(0)[0], sw, rtol=1e-13)  # total charge
    gam = np.sqrt(1 + (p ** 2).sum(0))
    lo, hi = gam.min(), gam.max()
    h = eng.spectrum(lo, hi, 50, "gamma")
    hw = np.histogram(gam, 50, (lo, hi), weights=w)[0]
    assert np.abs(h - hw).max() <= 1e-12 * np.abs(hw).max() + 2 * np.abs(w).max()  # a value on a bin edge may round either way
    assert np.isclose(h.sum(), hw.sum(), rtol=1e-12)
    hp = eng.spectrum(p[0].min(), p[0].max(), 33, "px")
    assert np.isclose(hp.sum(), sw, rtol=1e-12)
    eb = eng.download("EB")
    assert np.array_equal(eng.lineout("EB", 0, 0, 0), eb[:, 0, 0, 0])
    assert np.array_equal(eng.lineout("EB", 3, a["Mtot"] - 1, 4), eb[:, 3, a["Mtot"] - 1, 4])
    assert np.array_equal(eng.lineout("EG_fb", 2, 0, 5), eg[:, 2, 0, 5])
    eng.close()


def test_engine_static_kick_schedule(ofim, gfim):
    """'StaticKick' (BASELINE config 2, space-charge demo stage 1; chimera_main.py:106-125, 186): the field is
    rebuilt every step as the quasi-static field of the beam moving with its mean momentum (device reduction),
    rho deposited on coords_halfstep; resident engine against the reference sequence on the oracle."""
    from chimera_b200.engine import Engine

    S = SolverSetup(copy.deepcopy(SETUPS["static_m2"]))
    x, p, w = plasma(S, 2, 2, 91)
    p[0] += 50.0  # the demo's beam: px = 50 (doc/space-charge-demo cell 7)
    ref = RefRun(ofim, S, [RefSpecies(x, p, w)])
    eng = Engine(S)
    eng.add_species(x, p, w)
    ref.make_halfstep(px0=(50.0,))
    eng.make_halfstep(px0=(50.0,))
    compare_state(ref, eng, TOL, names=("J", "Rho", "EG_fb", "EB"))
    for _ in range(6):
        ref.make_step()
    eng.step(6)
    compare_state(ref, eng, 50 * TOL, names=("J", "Rho", "EG_fb", "EB"))
    eng.close()


@pytest.mark.parametrize("name,ions", [("real_m2", True), ("real_m3", False), ("env_m3", False)])
def test_step_graph_replay_matches_eager_steps(ofim, gfim, name, ions):
    """the fused step replayed as a CUDA graph between two re-binnings (csrc/engine.cu step_graphed) against the same
    engine stepping eagerly: 25 steps cross two re-binnings (graphs re-used after each) and both gradRho parities"""
    from chimera_b200.engine import Engine

    S = SolverSetup(copy.deepcopy(SETUPS[name]))
    x, p, w = plasma(S, 2, 2, 81)
    eg0 = seed_fields(S, 82, 0.1)
    engines = []
    for graph in (True, False):
        e = Engine(S, sort_every=9)
        e.set_graph(graph)
        e.add_species(x, p, w)
        if ions:
            xi, pi_, wi = plasma(S, 2, 2, 88)
            e.add_species(xi, 0 * pi_, -wi, charge=1.0, mass=1886.0, still=True)
        e.upload("EG_fb", eg0)
        e.make_halfstep(background=ions)
        e.step(25)
        engines.append(e)
    eng, eager = engines
    n, state = eng.graph_info()
    assert state == 1 and n == (2 if eng.cfg.space_charge else 1), (n, state)
    assert eager.graph_info()[0] == 0
    for nm in ("EG_fb", "EB", "J"):
        assert_close(eng.download(nm), eager.download(nm), 100 * TOL, nm)
    xa, _, pa, wa = eng.particles(0)
    xb, _, pb, wb = eager.particles(0)
    perm = match(wb, wa)
    assert_close(pa[:, perm], pb, 100 * TOL, "momenta")
    eng.close()
    eager.close()


@pytest.mark.parametrize("name,ions,und", [("real_m2", True, False), ("env_m1", False, True), ("static_m2", False, False)])
def test_one_step_calls_run_the_kernels_of_one_long_call(ofim, gfim, name, ions, und):
    """chimera_engine_step leaves the closing gather + push pending and the next call fuses it (csrc/engine.cu
    tail_pending): 14 one-step calls (with reads of the state in between, which must see completed momenta) against one
    14-step call on an engine that completes every call eagerly.  Crosses a re-binning; the time-dependent device of the
    envelope case checks the time the carried gather + push is evaluated at."""
    from chimera_b200.engine import Engine

    S = SolverSetup(copy.deepcopy(SETUPS[name]))
    x, p, w = plasma(S, 2, 2, 91)
    static = name.startswith("static")
    if static:
        p[0] += 50.0
    eg0 = seed_fields(S, 92, 0.1)
    engines = []
    for lazy in (True, False):
        e = Engine(S, sort_every=9)
        e.set_lazy_tail(lazy)
        e.set_graph(False)
        sid = e.add_species(x, p, w)
        if und:
            e.add_device("planewave", [0.3, 1.0, S.Args["leftX"], 40.0, 2.0, 0.1, 0.2], sid=sid)
        if ions:
            xi, pi_, wi = plasma(S, 2, 2, 98)
            e.add_species(xi, 0 * pi_, -wi, charge=1.0, mass=1886.0, still=True)
        if not static:
            e.upload("EG_fb", eg0)
        e.make_halfstep(**({"px0": (50.0,)} if static else {"background": ions}))
        engines.append(e)
    lazy, eager = engines
    seen = []
    for k in range(14):
        lazy.step(1)
        if k in (3, 10):  # a read in the middle completes the pending work: same momenta as the eager engine then
            seen.append(lazy.particles(0)[2].copy())
    eager.step(4)
    ref4 = eager.particles(0)[2].copy()
    eager.step(10)
    assert_close(seen[0], ref4, 100 * TOL, "momenta after 4 steps")
    for nm in ("EG_fb", "EB", "J"):
        # the envelope deposit sums terms that carry exp(-i kx0 x): a cancelling sum (compare_state), and the fused and
        # the separate kernels add them in different orders
        assert_close(lazy.download(nm), eager.download(nm), (20 if S.env and nm == "J" else 1) * 100 * TOL, nm)
    xa, _, pa, wa = lazy.particles(0)
    xb, _, pb, wb = eager.particles(0)
    perm = match(wb, wa)
    assert_close(pa[:, perm], pb, 100 * TOL, "momenta")
    assert_close(xa[:, perm], xb, 100 * TOL, "coords")
    lazy.close()
    eager.close()
